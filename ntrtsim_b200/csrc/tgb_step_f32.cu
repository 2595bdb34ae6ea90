/* Single-precision instantiation of the batched step (the throughput configuration). */
#define TGB_SCALAR float
#define TGB_OPS_NAME tgb_ops_f32
#include "tgb_inst.cuh"
