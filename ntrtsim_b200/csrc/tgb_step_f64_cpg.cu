/* Step kernels, double, controller class cpg. */
#define TGB_SCALAR double
#define TGB_STEP_CLASS tgb::TGB_CLASS_CPG
#define TGB_STEP_FN tgb_step_f64_cpg
#include "tgb_inst_step.cuh"
