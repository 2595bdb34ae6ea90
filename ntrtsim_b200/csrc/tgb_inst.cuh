/*
 * tgb_inst.cuh -- instantiates the I/O kernels for the scalar type TGB_SCALAR and defines the launch table TGB_OPS_NAME; the step
 * kernels of the three controller classes live in their own translation units (tgb_inst_step.cuh) and are declared here.
 * Included exactly once by tgb_step_f32.cu / tgb_step_f64.cu.
 */
#include <cstdlib>
#include "tgb_ops.h"
#include "tgb_step.cuh"

cudaError_t TGB_STEP_SIMPLE(const tgb_step_args& a, int lpe, int threads, cudaStream_t s, int* wide_ctas);
cudaError_t TGB_STEP_SINE(const tgb_step_args& a, int lpe, int threads, cudaStream_t s, int* wide_ctas);
cudaError_t TGB_STEP_CPG(const tgb_step_args& a, int lpe, int threads, cudaStream_t s, int* wide_ctas);

namespace {

typedef TGB_SCALAR scalar_t;

cudaError_t stepOrQuery(const tgb_step_args& a, int lpe, int threads, cudaStream_t s, int* wide_ctas) {
    switch (tgb::tgb_ctrl_class(a.ctrl_kind)) {
    case tgb::TGB_CLASS_CPG: return TGB_STEP_CPG(a, lpe, threads, s, wide_ctas);
    case tgb::TGB_CLASS_SINE: return TGB_STEP_SINE(a, lpe, threads, s, wide_ctas);
    default: return TGB_STEP_SIMPLE(a, lpe, threads, s, wide_ctas);
    }
}
cudaError_t step(const tgb_step_args& a, int lpe, int threads, cudaStream_t s) { return stepOrQuery(a, lpe, threads, s, nullptr); }
/* CTAs of the four-warp layout per SM for this controller class / lanes / ground; 0 when that layout is not compiled for them */
cudaError_t wide_ctas_per_sm(const tgb_step_args& a, int lpe, int* out) { return stepOrQuery(a, lpe, 128, nullptr, out); }

size_t step_smem(const tgb_step_args& a, int lpe, int threads) { return tgb::tgb_step_smem_bytes(a, lpe, threads, sizeof(scalar_t)); }

inline int blocksFor(int n, int threads) { return (n + threads - 1) / threads; }

cudaError_t pack_rods(const tgb::tgb_io_args& a, const double* src, int broadcast, cudaStream_t s) {
    const int n = a.env_count * a.R;
    if (n <= 0) return cudaSuccess;
    tgb::pack_rods_kernel<scalar_t><<<blocksFor(n, 256), 256, 0, s>>>(a, src, broadcast);
    return cudaGetLastError();
}
cudaError_t unpack_rods(const tgb::tgb_io_args& a, double* dst, cudaStream_t s) {
    const int n = a.env_count * a.R;
    if (n <= 0) return cudaSuccess;
    tgb::unpack_rods_kernel<scalar_t><<<blocksFor(n, 256), 256, 0, s>>>(a, dst);
    return cudaGetLastError();
}
cudaError_t pack_cables(const tgb::tgb_io_args& a, const double* src, int broadcast, int reset_motor, cudaStream_t s) {
    const int n = a.env_count * a.C;
    if (n <= 0) return cudaSuccess;
    tgb::pack_cables_kernel<scalar_t><<<blocksFor(n, 256), 256, 0, s>>>(a, src, broadcast, reset_motor);
    return cudaGetLastError();
}
cudaError_t cable_obs(const tgb::tgb_io_args& a, double* dst, cudaStream_t s) {
    const int n = a.env_count * a.C;
    if (n <= 0) return cudaSuccess;
    tgb::cable_obs_kernel<scalar_t><<<blocksFor(n, 256), 256, 0, s>>>(a, dst);
    return cudaGetLastError();
}
cudaError_t control_input(const tgb::tgb_io_args& a, const double* input, int broadcast, double dt, cudaStream_t s) {
    const int n = a.env_count * a.C;
    if (n <= 0) return cudaSuccess;
    tgb::control_input_kernel<scalar_t><<<blocksFor(n, 256), 256, 0, s>>>(a, input, broadcast, dt);
    return cudaGetLastError();
}
cudaError_t com(const tgb::tgb_io_args& a, int mode, double* scores, float* scores_f, cudaStream_t s) {
    if (a.env_count <= 0) return cudaSuccess;
    tgb::com_kernel<scalar_t><<<blocksFor(a.env_count, 128), 128, 0, s>>>(a, mode, scores, scores_f);
    return cudaGetLastError();
}

cudaError_t ccable_step(const tgb::tgb_ccable_args& a, cudaStream_t s) {
    if (a.env_count <= 0 || a.n_cc <= 0) return cudaSuccess;
    /* One thread per environment on a data-dependent path: the fewer environments share a warp, the less their paths serialise.
       TGB_CC_BLOCK environments per CTA (= per warp), default 8.  Measured on B200 (4096 environments): 32 / 8 / 4 / 2 / 1 give 4.1 / 4.4 / 4.2 / 4.0 / 3.7e7 env-steps/s -- a launch lasts as long as its slowest environment's own dependent chain, not as long as a warp's divergence. */
    static const int block = [] { const char* e = std::getenv("TGB_CC_BLOCK"); const int v = e ? std::atoi(e) : 0; return v >= 1 && v <= 32 ? v : 8; }();
    tgb::ccable_step_kernel<scalar_t><<<blocksFor(a.env_count, block), block, 0, s>>>(a);
    return cudaGetLastError();
}
cudaError_t ccable_obs(const tgb::tgb_ccable_args& a, double* obs, int cable, double* anchors, cudaStream_t s) {
    const int n = a.env_count * a.n_cc;
    if (n <= 0) return cudaSuccess;
    tgb::ccable_obs_kernel<scalar_t><<<blocksFor(n, 128), 128, 0, s>>>(a, obs, cable, anchors);
    return cudaGetLastError();
}
cudaError_t ccable_control(const tgb::tgb_ccable_args& a, const double* input, int broadcast, cudaStream_t s) {
    const int n = a.env_count * a.n_cc;
    if (n <= 0) return cudaSuccess;
    tgb::ccable_control_kernel<scalar_t><<<blocksFor(n, 64), 64, 0, s>>>(a, input, broadcast);
    return cudaGetLastError();
}

const tgb_ops kOps = {sizeof(scalar_t), step, wide_ctas_per_sm, step_smem, pack_rods, unpack_rods, pack_cables, cable_obs, control_input, com,
                      ccable_step, ccable_obs, ccable_control};

} // namespace

const tgb_ops* TGB_OPS_NAME() { return &kOps; }
