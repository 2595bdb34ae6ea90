/* Step kernels, double, controller class simple. */
#define TGB_SCALAR double
#define TGB_STEP_CLASS tgb::TGB_CLASS_SIMPLE
#define TGB_STEP_FN tgb_step_f64_simple
#include "tgb_inst_step.cuh"
