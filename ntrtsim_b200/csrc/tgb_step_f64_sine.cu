/* Step kernels, double, controller class sine. */
#define TGB_SCALAR double
#define TGB_STEP_CLASS tgb::TGB_CLASS_SINE
#define TGB_STEP_FN tgb_step_f64_sine
#include "tgb_inst_step.cuh"
