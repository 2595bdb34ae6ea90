/* Step kernels, float, controller class simple. */
#define TGB_SCALAR float
#define TGB_STEP_CLASS tgb::TGB_CLASS_SIMPLE
#define TGB_STEP_FN tgb_step_f32_simple
#include "tgb_inst_step.cuh"
