/* Step kernels, float, controller class sine. */
#define TGB_SCALAR float
#define TGB_STEP_CLASS tgb::TGB_CLASS_SINE
#define TGB_STEP_FN tgb_step_f32_sine
#include "tgb_inst_step.cuh"
