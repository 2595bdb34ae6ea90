/* Double-precision instantiation of the batched step; built with --fmad=false (see tgb_ops.h). */
#define TGB_SCALAR double
#define TGB_OPS_NAME tgb_ops_f64
#include "tgb_inst.cuh"
