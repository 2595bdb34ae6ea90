/*
 * tgb_inst_step.cuh -- instantiates tgb_step_kernel for the scalar type TGB_SCALAR and the controller class TGB_STEP_CLASS
 * (all lanes-per-env values, with and without ground) and defines the launch function TGB_STEP_FN.  One translation unit per
 * (scalar type, class) keeps the build parallel; see the Makefile.
 */
#include "tgb_ops.h"
#include "tgb_step.cuh"

#include <mutex>

namespace {

typedef TGB_SCALAR scalar_t;

/* cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute of a kernel: remember the largest size set
   so far per device ordinal, never lower it (a smaller world must not take the room away from a larger one created earlier), and
   serialise the bookkeeping (worlds of different host threads share these kernels). */
struct SmemAttrCache {
    enum { kMaxDevices = 64 };
    size_t configured[kMaxDevices] = {};
    std::mutex lock;
    template <typename Kernel>
    cudaError_t ensure(Kernel kernel, size_t smem) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= kMaxDevices) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        std::lock_guard<std::mutex> g(lock);
        if (smem > configured[dev]) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[dev] = smem;
        }
        return cudaSuccess;
    }
};

template <int LPE, bool GROUND, bool OBST, bool WARM>
cudaError_t launch_inst(const tgb_step_args& a, int threads, cudaStream_t s) {
    const size_t smem = tgb::tgb_step_smem_bytes(a, LPE, threads, sizeof(scalar_t));
    static SmemAttrCache cache;
    {
        cudaError_t e = cache.ensure(tgb::tgb_step_kernel<scalar_t, LPE, TGB_STEP_CLASS, GROUND, 0, OBST, WARM>, smem);
        if (e != cudaSuccess) return e;
    }
    const int groupsPerBlock = threads / LPE;
    const int slots = a.env_list != nullptr ? a.env_list_slots : a.n_env - a.env_begin;
    const int grid = (slots + groupsPerBlock - 1) / groupsPerBlock;
    if (grid <= 0) return cudaSuccess;
    tgb::tgb_step_kernel<scalar_t, LPE, TGB_STEP_CLASS, GROUND, 0, OBST, WARM><<<grid, threads, smem, s>>>(a);
    return cudaGetLastError();
}

/* the persistent-manifold instantiation only when the world asks for warm starting (tgb_step_args.warm) */
template <int LPE, bool GROUND, bool OBST = false>
cudaError_t launch_impl(const tgb_step_args& a, int threads, cudaStream_t s) {
    if (a.warm != nullptr) return launch_inst<LPE, GROUND, OBST, true>(a, threads, s);
    return launch_inst<LPE, GROUND, OBST, false>(a, threads, s);
}

/* one cache per instantiation of the four-warp layout, shared by its launch and its occupancy query */
template <int LPE, bool GROUND, int NW>
SmemAttrCache& wideCache() {
    static SmemAttrCache cache;
    return cache;
}

/* the four-warp layout: 128 threads, 32 / LPE environments per CTA */
template <int LPE, bool GROUND, int NW>
cudaError_t launch_wide(const tgb_step_args& a, cudaStream_t s) {
    const size_t smem = tgb::tgb_step_smem_bytes(a, LPE, 128, sizeof(scalar_t));
    {
        cudaError_t e = wideCache<LPE, GROUND, NW>().ensure(tgb::tgb_step_kernel<scalar_t, LPE, TGB_STEP_CLASS, GROUND, NW>, smem);
        if (e != cudaSuccess) return e;
    }
    const int groupsPerBlock = 32 / LPE;
    const int grid = (a.n_env - a.env_begin + groupsPerBlock - 1) / groupsPerBlock;
    if (grid <= 0) return cudaSuccess;
    tgb::tgb_step_kernel<scalar_t, LPE, TGB_STEP_CLASS, GROUND, NW><<<grid, 32 * NW, smem, s>>>(a);
    return cudaGetLastError();
}

/* CTAs of the four-warp layout one SM holds (registers and shared memory decide) */
template <int LPE, bool GROUND>
cudaError_t wide_ctas_per_sm(const tgb_step_args& a, int* out) {
    tgb_step_args aw = a;
    aw.wide = 1;
    const size_t smem = tgb::tgb_step_smem_bytes(aw, LPE, 128, sizeof(scalar_t));
    cudaError_t e = wideCache<LPE, GROUND, 4>().ensure(tgb::tgb_step_kernel<scalar_t, LPE, TGB_STEP_CLASS, GROUND, 4>, smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, tgb::tgb_step_kernel<scalar_t, LPE, TGB_STEP_CLASS, GROUND, 4>, 128, smem);
}

template <int LPE>
cudaError_t launch(const tgb_step_args& a, int threads, cudaStream_t s, int* wide_ctas) {
    if (wide_ctas) {
        *wide_ctas = 0;
        if constexpr (TGB_STEP_CLASS != tgb::TGB_CLASS_CPG && (LPE == 4 || LPE == 8)) {
            if (a.ground_kind == TGB_GROUND_EMPTY) return wide_ctas_per_sm<LPE, false>(a, wide_ctas);
            return wide_ctas_per_sm<LPE, true>(a, wide_ctas);
        }
        return cudaSuccess;
    }
    if (a.wide && a.env_list != nullptr) return cudaErrorInvalidValue; /* lists of environments: narrow layouts only */
    if constexpr (TGB_STEP_CLASS != tgb::TGB_CLASS_CPG && (LPE == 4 || LPE == 8)) {
        if (a.wide) {
            if (a.ground_kind == TGB_GROUND_EMPTY) return launch_wide<LPE, false, 4>(a, s);
            return launch_wide<LPE, true, 4>(a, s);
        }
    }
    if (a.wide) return cudaErrorInvalidValue;
    if (a.n_obst > 0) {
        /* worlds with static obstacles: own instantiation (rows with normals); not compiled for the CPG class */
        if constexpr (TGB_STEP_CLASS != tgb::TGB_CLASS_CPG) return launch_impl<LPE, true, true>(a, threads, s);
        return cudaErrorNotSupported;
    }
    if (a.ground_kind == TGB_GROUND_EMPTY) return launch_impl<LPE, false>(a, threads, s);
    return launch_impl<LPE, true>(a, threads, s);
}

} // namespace

cudaError_t TGB_STEP_FN(const tgb_step_args& a, int lpe, int threads, cudaStream_t s, int* wide_ctas) {
    if (threads < 32 || threads > 128 || (threads % 32) != 0) return cudaErrorInvalidValue;
    switch (lpe) {
    case 2: return launch<2>(a, threads, s, wide_ctas);
    case 4: return launch<4>(a, threads, s, wide_ctas);
    case 8: return launch<8>(a, threads, s, wide_ctas);
    case 16: return launch<16>(a, threads, s, wide_ctas);
    case 32: return launch<32>(a, threads, s, wide_ctas);
    default: return cudaErrorInvalidValue;
    }
}
