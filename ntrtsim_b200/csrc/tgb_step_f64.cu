/* Double-precision launch table and I/O kernels; the step kernels are in tgb_step_f64_{simple,sine,cpg}.cu (built with --fmad=false, see tgb_ops.h) */
#define TGB_SCALAR double
#define TGB_OPS_NAME tgb_ops_f64
#define TGB_STEP_SIMPLE tgb_step_f64_simple
#define TGB_STEP_SINE tgb_step_f64_sine
#define TGB_STEP_CPG tgb_step_f64_cpg
#include "tgb_inst.cuh"
