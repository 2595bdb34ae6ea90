/* Single-precision launch table and I/O kernels; the step kernels are in tgb_step_f32_{simple,sine,cpg}.cu */
#define TGB_SCALAR float
#define TGB_OPS_NAME tgb_ops_f32
#define TGB_STEP_SIMPLE tgb_step_f32_simple
#define TGB_STEP_SINE tgb_step_f32_sine
#define TGB_STEP_CPG tgb_step_f32_cpg
#include "tgb_inst.cuh"
