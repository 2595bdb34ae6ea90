/* Step kernels, float, controller class cpg. */
#define TGB_SCALAR float
#define TGB_STEP_CLASS tgb::TGB_CLASS_CPG
#define TGB_STEP_FN tgb_step_f32_cpg
#include "tgb_inst_step.cuh"
