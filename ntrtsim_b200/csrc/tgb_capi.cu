/*
 * tgb_capi.cu -- the C ABI of include/ntrtsim_b200.h over the CUDA kernels.  Each entry point cites the reference
 * interface it replaces in the header; this file owns the device buffers, the packing of the flattened model and the
 * launch geometry.  There is no CPU implementation behind any call: without a usable CUDA device tgb_create fails.
 */
#include "../../include/ntrtsim_b200.h"
#include "tgb_ops.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;

tgb_status fail(tgb_status st, const std::string& msg) {
    g_error = msg;
    return st;
}
tgb_status cudaFail(cudaError_t e, const char* what) {
    g_error = std::string(what) + ": " + cudaGetErrorString(e);
    return TGB_ERR_CUDA;
}
#define TGB_CUDA(call)                                       \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return cudaFail(e__, #call); \
    } while (0)

int nextPow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

} // namespace

struct tgb_world {
    tgb_world_cfg cfg;
    tgb_ground_desc ground;
    std::vector<tgb_body_desc> bodies;
    std::vector<tgb_shape_desc> shapes;
    std::vector<tgb_cable_desc> cables;
    std::vector<tgb_contact_cable_desc> contactCables; /* tgb_add_contact_cables; body indices already absolute */
    int lastModelBody0 = 0;
    double* d_cc_state = nullptr;  /* [n_env][nCC][TGB_CC_STATE] (tgb_ccable.cuh) */
    double* d_cc_model = nullptr;  /* [nCC][TGB_CC_MODEL] */
    /* CUDA graphs of `steps` x (step kernel, contact-cable kernel) for one range of environments, replayed by tgb_step: a world with
       contact cables is launch bound (two small launches per simulation step) */
    struct CcGraph { tgb_step_args args; int steps, envBegin, envCount; cudaGraphExec_t exec; };
    std::vector<CcGraph> ccGraphs;
    bool ccLaunchedDirectly = false; /* the kernels have been launched once outside a capture (function attributes set) */
    bool finalized = false;
    const tgb_ops* ops = nullptr;
    size_t scalarBytes = 4;
    int R = 0, C = 0, S = 0, stride = 0, capContacts = 0, frameBody = 0, nInc = 0, maxInc = 0, nObst = 0, modelObstOff = 0;
    int nPairs = 0, modelPairOff = 0, pairCap = TGB_DEFAULT_PAIR_CONTACT_CAP; /* contact between bodies (tgb_pairs.cuh) */
    float* d_min_clearance = nullptr;
    void* d_warm = nullptr;        /* persistent manifolds between launches: T[n_env][warmWords] (tgb_dev.h TGB_WARM_*) */
    int warmWords = 0;
    double warmFactor = TGB_DEFAULT_WARM_STARTING; /* btContactSolverInfo::m_warmstartingFactor */
    int lanes = 0, threads = 0, lanesOverride = 0, threadsOverride = 0, wideMode = -1;
    mutable int wideCtasPerSm[2] = {-1, -1}; /* by controller class (simple, sine); -1 = not asked yet */
    int smCount = 148;
    cudaStream_t stream = nullptr;
    void* d_state = nullptr;
    double* d_frame = nullptr;
    double* d_ctrl = nullptr;
    void* d_max_tension = nullptr; /* tgb_track_max_tension */
    double* d_sine = nullptr;   /* sine controller phase pairs, see tgb_step_args::sine_state */
    double* d_com0 = nullptr;
    int* d_status = nullptr;
    void* d_model_t = nullptr;
    int* d_model_i = nullptr;
    int modelTWords = 0, modelIWords = 0, modelCrecOff = 0, modelCirecOff = 0;
    double* d_stage = nullptr;  /* staging for double records exchanged with the host */
    size_t stageDoubles = 0;
    double* h_stage = nullptr;  /* pinned */
    size_t hStageDoubles = 0;
    float* d_scores = nullptr;
    int ctrlKind = TGB_CTRL_NONE, ctrlStride = 0;
    /* CPG + impedance controller */
    tgb_cpg_desc cpg;
    int* d_cpg_topo = nullptr;
    double* d_cpg_imp = nullptr;
    double* d_cpg_state = nullptr;
    double time = 0.0;
    int64_t launches = 0;
    double heightLo = 1.0, heightHi = 0.0; /* disabled */
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned long long* d_stats = nullptr; /* TGB_STATS=1 with a -DTGB_STATS build of the kernels (tools/) */
    /* tail split (launchSplit below): steps per environment with rows between two bodies in its last launch, and per range of
       environments in flight (slot 0: the whole world on `stream`; slot 1 + k: chunk k of tgb_rollout_host) the two lists of a launch */
    enum { kSplitSlots = 5 };
    int* d_coupled = nullptr;       /* [n_env] */
    int* d_split_count = nullptr;   /* [kSplitSlots][4]: slow candidates, fast environments, medium candidates, spare */
    int* d_split_slow = nullptr;    /* [kSplitSlots][splitSlowCap] */
    int* d_split_medium = nullptr;  /* [kSplitSlots][splitMediumCap]: environments run two to a warp (16 lanes each) */
    int splitMediumCap = 0;
    cudaStream_t splitStream2[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t splitJoin2[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int* d_split_fast = nullptr;    /* [n_env]: a range's fast list sits at its env_begin */
    int splitSlowCap = 0, splitMode = -1 /* -1 automatic, 0 off, 1 forced (tests) */, coupledLaunchSteps = 0;
    cudaStream_t splitStream[kSplitSlots] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t splitFork[kSplitSlots] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t splitJoin[kSplitSlots] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    /* tgb_rollout_host pipelines env chunks over these streams: the copies of one chunk overlap the steps of the others */
    enum { kPipe = 4 };
    cudaStream_t pipe[kPipe] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pipeDone[kPipe] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pipeStart = nullptr;

    tgb::tgb_io_args ioArgs(int envBegin, int envCount) const {
        tgb::tgb_io_args a;
        a.state = d_state; a.frame = d_frame; a.com0 = d_com0; a.model_t = d_model_t; a.model_i = d_model_i;
        a.n_env = cfg.n_env; a.R = R; a.C = C; a.S = S; a.stride = stride; a.frame_body = frameBody;
        a.frame_static = bodies[frameBody].mass == 0.0 ? 1 : 0;
        a.score_body = ctrlKind == TGB_CTRL_CPG_IMPEDANCE ? cpg.score_body : -1;
        a.status = d_status;
        a.env_begin = envBegin; a.env_count = envCount;
        return a;
    }
    tgb::tgb_ccable_args ccArgs(int envBegin, int envCount, double dt) const {
        tgb::tgb_ccable_args a;
        a.state = d_state; a.frame = d_frame; a.model_t = d_model_t; a.model_i = d_model_i;
        a.cc_state = d_cc_state; a.cc_model = d_cc_model;
        a.n_env = cfg.n_env; a.env_begin = envBegin; a.env_count = envCount; a.R = R; a.S = S; a.stride = stride;
        a.n_cc = (int)contactCables.size(); a.dt = dt;
        return a;
    }
    tgb_status ensureStage(size_t doubles) {
        if (doubles > stageDoubles) {
            if (d_stage) cudaFree(d_stage);
            d_stage = nullptr; stageDoubles = 0;
            TGB_CUDA(cudaMalloc(&d_stage, doubles * sizeof(double)));
            stageDoubles = doubles;
        }
        return TGB_OK;
    }
    tgb_status ensureHostStage(size_t doubles) {
        if (doubles > hStageDoubles) {
            if (h_stage) cudaFreeHost(h_stage);
            h_stage = nullptr; hStageDoubles = 0;
            TGB_CUDA(cudaMallocHost(&h_stage, doubles * sizeof(double)));
            hStageDoubles = doubles;
        }
        return TGB_OK;
    }
    void release() {
        if (d_state) cudaFree(d_state);
        if (d_frame) cudaFree(d_frame);
        if (d_ctrl) cudaFree(d_ctrl);
        if (d_sine) cudaFree(d_sine);
        if (d_max_tension) cudaFree(d_max_tension);
        if (d_com0) cudaFree(d_com0);
        if (d_status) cudaFree(d_status);
        if (d_model_t) cudaFree(d_model_t);
        if (d_model_i) cudaFree(d_model_i);
        if (d_stage) cudaFree(d_stage);
        if (h_stage) cudaFreeHost(h_stage);
        if (d_scores) cudaFree(d_scores);
        if (d_min_clearance) cudaFree(d_min_clearance);
        if (d_coupled) cudaFree(d_coupled);
        if (d_split_count) cudaFree(d_split_count);
        if (d_split_slow) cudaFree(d_split_slow);
        if (d_split_medium) cudaFree(d_split_medium);
        for (int i = 0; i < 5; ++i) {
            if (splitJoin2[i]) cudaEventDestroy(splitJoin2[i]);
            if (splitStream2[i]) cudaStreamDestroy(splitStream2[i]);
        }
        if (d_split_fast) cudaFree(d_split_fast);
        for (int i = 0; i < kSplitSlots; ++i) {
            if (splitStream[i]) cudaStreamDestroy(splitStream[i]);
            if (splitFork[i]) cudaEventDestroy(splitFork[i]);
            if (splitJoin[i]) cudaEventDestroy(splitJoin[i]);
        }
        if (d_warm) cudaFree(d_warm);
        if (d_cpg_topo) cudaFree(d_cpg_topo);
        if (d_cpg_imp) cudaFree(d_cpg_imp);
        if (d_cpg_state) cudaFree(d_cpg_state);
        for (CcGraph& g : ccGraphs) cudaGraphExecDestroy(g.exec);
        if (d_cc_state) cudaFree(d_cc_state);
        if (d_cc_model) cudaFree(d_cc_model);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        for (int i = 0; i < kPipe; ++i) {
            if (pipeDone[i]) cudaEventDestroy(pipeDone[i]);
            if (pipe[i]) cudaStreamDestroy(pipe[i]);
        }
        if (pipeStart) cudaEventDestroy(pipeStart);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

bool rangeOk(const tgb_world* w, int envBegin, int envCount) {
    return envBegin >= 0 && envCount >= 0 && (int64_t)envBegin + envCount <= w->cfg.n_env;
}

template <typename T>
void packModel(const tgb_world* w, std::vector<unsigned char>& tbytes, std::vector<int>& ints, int& nInc, int& crecOff, int& cirecOff, int& maxInc,
               int& nObst, int& obstOff, int& nPairs, int& pairOff) {
    const int R = w->R, C = w->C, S = w->S;
    crecOff = (TGB_NBF * R + TGB_NSF * S + TGB_NCF * C + 3) & ~3;
    std::vector<T> t((size_t)crecOff + (size_t)TGB_CREC * C, T(0));
    T* bt = t.data();
    T* st = bt + (size_t)TGB_NBF * R;
    T* ct = st + (size_t)TGB_NSF * S;
    std::vector<std::vector<int> > incidence(R);
    for (int c = 0; c < C; ++c) {
        incidence[w->cables[c].body_a].push_back(c);
        incidence[w->cables[c].body_b].push_back(c | (1 << 24));
    }
    nInc = 2 * C;
    cirecOff = (TGB_NBI * R + TGB_NCI * C + nInc + 3) & ~3;
    ints.assign((size_t)cirecOff + (size_t)TGB_CIREC * C, 0);
    int* bi = ints.data();
    int* ci = bi + (size_t)TGB_NBI * R;
    int* inc = ci + (size_t)TGB_NCI * C;
    int incPos = 0;
    maxInc = 0;
    std::vector<int> slotA(std::max(1, C), 0), slotB(std::max(1, C), 0);
    for (int b = 0; b < R; ++b) {
        const tgb_body_desc& d = w->bodies[b];
        /* btRigidBody::setMassProps: inverse mass and inverse local inertia; static when mass == 0 */
        const double invMass = d.mass != 0.0 ? 1.0 / d.mass : 0.0;
        double invI[3];
        for (int k = 0; k < 3; ++k) invI[k] = (d.mass != 0.0 && d.inertia[k] != 0.0) ? 1.0 / d.inertia[k] : 0.0;
        bt[BF_INVMASS * R + b] = (T)invMass;
        /* btRigidBody::setGravity: m_gravity = acceleration * (1 / inverseMass) */
        bt[BF_GFY * R + b] = invMass != 0.0 ? (T)(-w->cfg.gravity * (1.0 / invMass)) : T(0);
        for (int k = 0; k < 3; ++k) {
            bt[(BF_INVIX + k) * R + b] = (T)invI[k];
            bt[(BF_IX + k) * R + b] = invI[k] != 0.0 ? (T)(1.0 / invI[k]) : T(0);
        }
        double mu = d.friction * w->ground.friction;
        mu = std::max(-10.0, std::min(10.0, mu));
        bt[BF_MU * R + b] = (T)mu;
        bt[BF_RESTITUTION * R + b] = (T)d.restitution;
        bt[BF_MASS * R + b] = (T)d.mass;
        bt[BF_FRICTION * R + b] = (T)d.friction;
        {
            /* every capsule of the body (tgb_pairs.cuh) lies within this radius of the body origin */
            double bound = 0.0;
            for (int sI = d.shape_begin; sI < d.shape_begin + d.shape_count; ++sI) {
                const tgb_shape_desc& sh = w->shapes[sI];
                if (sh.kind == TGB_SHAPE_BOX) continue;
                const double segH = sh.kind == TGB_SHAPE_SPHERE ? 0.0 : std::max(0.0, sh.half_length - sh.radius);
                bound = std::max(bound, std::sqrt(sh.pos[0] * sh.pos[0] + sh.pos[1] * sh.pos[1] + sh.pos[2] * sh.pos[2]) + segH + sh.radius);
            }
            bt[BF_BOUND * R + b] = (T)(bound * (1.0 + 1e-6));
            /* a plain rod / ball: one shape that is not a box, sitting in the body frame */
            bt[BF_SEGH * R + b] = T(-1); bt[BF_SEGR * R + b] = T(0);
            if (d.shape_count == 1) {
                const tgb_shape_desc& sh = w->shapes[d.shape_begin];
                const bool aligned = sh.pos[0] == 0.0 && sh.pos[1] == 0.0 && sh.pos[2] == 0.0 && sh.quat[0] == 0.0 && sh.quat[1] == 0.0 &&
                                     sh.quat[2] == 0.0 && sh.quat[3] != 0.0;
                if (aligned && sh.kind != TGB_SHAPE_BOX) {
                    bt[BF_SEGH * R + b] = sh.kind == TGB_SHAPE_SPHERE ? T(0) : (T)std::max(0.0, sh.half_length - sh.radius);
                    bt[BF_SEGR * R + b] = (T)sh.radius;
                }
            }
        }
        bi[BI_SHAPE_BEGIN * R + b] = d.shape_begin;
        bi[BI_SHAPE_COUNT * R + b] = d.shape_count;
        bi[BI_INC_BEGIN * R + b] = incPos;
        bi[BI_INC_COUNT * R + b] = (int)incidence[b].size();
        bi[BI_GYRO * R + b] = d.gyroscopic;
        /* applyImpulse order inside one tgSimulation::step is cable order, end A before end B */
        std::vector<int> sorted = incidence[b];
        std::sort(sorted.begin(), sorted.end(), [](int x, int y) {
            const int cx = x & 0xffffff, cy = y & 0xffffff;
            return cx != cy ? cx < cy : x < y;
        });
        for (size_t k = 0; k < sorted.size(); ++k) {
            inc[incPos++] = sorted[k];
            /* position of this cable end in the body's list, for the cable's record below */
            const int cab = sorted[k] & 0xffffff;
            (((sorted[k] >> 24) != 0) ? slotB : slotA)[cab] = (int)k;
        }
        maxInc = std::max(maxInc, (int)sorted.size());
    }
    for (int s = 0; s < S; ++s) {
        const tgb_shape_desc& d = w->shapes[s];
        st[SF_RADIUS * S + s] = (T)d.radius;
        st[SF_HALFLEN * S + s] = (T)d.half_length;
        st[SF_MARGIN * S + s] = (T)d.margin;
        st[SF_POSX * S + s] = (T)d.pos[0]; st[SF_POSY * S + s] = (T)d.pos[1]; st[SF_POSZ * S + s] = (T)d.pos[2];
        /* btMatrix3x3::setRotation in double, then narrowed */
        const double x = d.quat[0], y = d.quat[1], z = d.quat[2], q = d.quat[3];
        const double dd = x * x + y * y + z * z + q * q;
        const double sc = 2.0 / dd;
        const double xs = x * sc, ys = y * sc, zs = z * sc;
        const double wx = q * xs, wy = q * ys, wz = q * zs, xx = x * xs, xy = x * ys, xz = x * zs, yy = y * ys, yz = y * zs, zz = z * zs;
        const double m[9] = {1.0 - (yy + zz), xy - wz, xz + wy, xy + wz, 1.0 - (xx + zz), yz - wx, xz - wy, yz + wx, 1.0 - (xx + yy)};
        for (int k = 0; k < 9; ++k) st[(SF_ROT0 + k) * S + s] = (T)m[k];
        st[SF_KIND * S + s] = (T)d.kind;
        st[SF_HALFZ * S + s] = (T)d.half_extent_z;
        st[SF_SEGH * S + s] = d.kind == TGB_SHAPE_CYLINDER ? (T)std::max(0.0, d.half_length - d.radius) : T(0);
    }
    for (int c = 0; c < C; ++c) {
        const tgb_cable_desc& d = w->cables[c];
        for (int k = 0; k < 3; ++k) { ct[(CF_LAX + k) * C + c] = (T)d.local_a[k]; ct[(CF_LBX + k) * C + c] = (T)d.local_b[k]; }
        ct[CF_K * C + c] = (T)d.stiffness; ct[CF_C * C + c] = (T)d.damping;
        ct[CF_MAXTENS * C + c] = (T)d.max_tension; ct[CF_TVEL * C + c] = (T)d.target_velocity;
        ct[CF_MINAL * C + c] = (T)d.min_actual_length; ct[CF_MINRL * C + c] = (T)d.min_rest_length;
        ct[CF_MAXSTRETCH * C + c] = (T)(d.max_tension / d.stiffness);
        ct[CF_MRADIUS * C + c] = (T)d.motor_radius; ct[CF_MFRICTION * C + c] = (T)d.motor_friction; ct[CF_MINERTIA * C + c] = (T)d.motor_inertia;
        ci[CI_A * C + c] = d.body_a; ci[CI_B * C + c] = d.body_b;
        ci[CI_ACTUATOR * C + c] = d.actuator; ci[CI_BACKDRIVABLE * C + c] = d.backdrivable;
        /* the same constants as one record per cable (TGB_CREC / TGB_CIREC) */
        T* r = t.data() + crecOff + (size_t)TGB_CREC * c;
        for (int k = 0; k < 3; ++k) { r[k] = (T)d.local_a[k]; r[4 + k] = (T)d.local_b[k]; }
        r[3] = (T)d.stiffness; r[7] = (T)d.damping;
        r[8] = (T)d.target_velocity; r[9] = (T)d.max_tension; r[10] = (T)(d.max_tension / d.stiffness); r[11] = (T)d.min_actual_length;
        r[12] = (T)d.min_rest_length;
        int* ri = ints.data() + cirecOff + (size_t)TGB_CIREC * c;
        ri[0] = d.body_a; ri[1] = d.body_b; ri[2] = d.actuator | (d.backdrivable ? 1 << 8 : 0); ri[3] = slotA[c] | (slotB[c] << 16);
    }
    /* static obstacles: the box shapes of static bodies (tgBoxInfo with density 0, src/models/obstacles) */
    obstOff = (int)ints.size();
    nObst = 0;
    for (int b = 0; b < R; ++b) {
        if (w->bodies[b].mass != 0.0) continue;
        for (int s = w->bodies[b].shape_begin; s < w->bodies[b].shape_begin + w->bodies[b].shape_count; ++s)
            if (w->shapes[s].kind == TGB_SHAPE_BOX) { ints.push_back(b); ints.push_back(s); ++nObst; }
    }
    /* pairs of bodies that can collide: at least one of them dynamic, both with a shape that is not a box */
    pairOff = (int)ints.size();
    nPairs = 0;
    {
        std::vector<char> hasCapsule(R, 0);
        for (int b = 0; b < R; ++b)
            for (int sI = w->bodies[b].shape_begin; sI < w->bodies[b].shape_begin + w->bodies[b].shape_count; ++sI)
                if (w->shapes[sI].kind != TGB_SHAPE_BOX) hasCapsule[b] = 1;
        for (int ia = 0; ia < R; ++ia)
            for (int ib = ia + 1; ib < R; ++ib) {
                if (w->bodies[ia].mass == 0.0 && w->bodies[ib].mass == 0.0) continue;
                if (!hasCapsule[ia] || !hasCapsule[ib]) continue;
                ints.push_back(ia + 32 * ib);
                ++nPairs;
            }
    }
    tbytes.resize(t.size() * sizeof(T));
    std::memcpy(tbytes.data(), t.data(), tbytes.size());
}

/* The four-warp layout of the step kernel (tgb_step.cuh, WIDE) shortens a step by ~28 % (an environment's cables run in one round)
   but holds four times the threads per environment: it pays while all its CTAs are resident at once (registers decide: 3 CTAs = 12
   SUPERball environments per SM, ~1800 per B200; measured at 4096 environments: 5.8e8 env-steps/s against 9.6e8, the grid then runs
   in 2.3 waves).  Automatic choice: wide when the whole grid is resident.  tgb_set_wide_layout / TGB_WIDE=0|1 override. */
bool wideLayout(const tgb_world* w) {
    if (w->ctrlKind == TGB_CTRL_CPG_IMPEDANCE || w->nObst > 0) return false;
    if (w->warmFactor > 0.0) return false; /* the persistent manifolds are compiled for the narrow layout only */
    if (w->lanes != 4 && w->lanes != 8) return false;
    if (w->wideMode >= 0) return w->wideMode != 0;
    static const char* envWide = std::getenv("TGB_WIDE");
    if (envWide) return std::atoi(envWide) != 0;
    const int cls = w->ctrlKind == TGB_CTRL_ESCAPE_SINE ? 1 : 0;
    if (w->wideCtasPerSm[cls] < 0) {
        tgb_step_args a;
        std::memset(&a, 0, sizeof(a));
        a.R = w->R; a.C = w->C; a.S = w->S; a.max_inc = w->maxInc; a.stride = w->stride; a.cap_contacts = w->capContacts;
        a.ctrl_kind = w->ctrlKind; a.ctrl_stride = std::max(w->C, TGB_ESCAPE_PARAMS); a.ground_kind = w->ground.kind;
        a.model_t_words = w->modelTWords; a.model_i_words = w->modelIWords;
        a.n_pairs = w->pairCap != 0 ? w->nPairs : 0; a.pair_cap = w->pairCap;
        a.warm = w->warmFactor > 0.0 ? w->d_warm : nullptr; a.n_dyn = 0;
        for (int b = 0; b < w->R; ++b) a.n_dyn += w->bodies[b].mass != 0.0;
        int ctas = 0;
        if (w->ops->wide_ctas_per_sm(a, w->lanes, &ctas) != cudaSuccess) { cudaGetLastError(); ctas = 0; }
        w->wideCtasPerSm[cls] = ctas;
    }
    /* with contact between bodies compiled in, the four-warp kernel no longer pays (its 168 registers spill 632 bytes; capped at two CTAs
       per SM instead it has 255 and still loses): 256 / 512 / 1024 / 1184 environments 3.9e7 / 7.8e7 / 1.49e8 / 1.73e8 against
       3.8e7 / 1.12e8 / 1.98e8 / 2.27e8 for the narrow layout (round 2, same box).  Automatic choice: narrow; tgb_set_wide_layout(1) still forces it */
    if (w->nPairs > 0 && w->pairCap != 0) return false;
    /* without ground the narrow layout already keeps three cables per lane in flight: no gain measured (3-prism, 256..1024 envs) */
    if (w->ground.kind == TGB_GROUND_EMPTY) return false;
    const int envsPerCta = 32 / w->lanes;
    const long long ctas = ((long long)w->cfg.n_env + envsPerCta - 1) / envsPerCta;
    return ctas <= (long long)w->wideCtasPerSm[cls] * w->smCount;
}

/* thresholds of the tail split's classification (environment variables: tuning aids) */
struct SplitTuning { int totalDiv, recentThreshold, mediumDiv, mediumCap; };
const SplitTuning& splitTuning() {
    static const SplitTuning t = [] {
        SplitTuning v = {8, 8, 64, 0}; /* the middle class is off: measured below */
        if (const char* e = std::getenv("TGB_SPLIT_MEDIUM_DIV")) v.mediumDiv = std::max(1, std::atoi(e));
        if (const char* e = std::getenv("TGB_SPLIT_MEDIUM_CAP")) v.mediumCap = std::max(0, std::atoi(e));
        if (const char* e = std::getenv("TGB_SPLIT_TOTAL_DIV")) v.totalDiv = std::max(1, std::atoi(e));
        if (const char* e = std::getenv("TGB_SPLIT_RECENT_THR")) v.recentThreshold = std::max(1, std::atoi(e));
        return v;
    }();
    return t;
}

tgb_step_args stepArgs(const tgb_world* w, double dt, int nSteps) {
    tgb_step_args a;
    std::memset(&a, 0, sizeof(a));
    a.state = w->d_state; a.frame = w->d_frame; a.ctrl = w->ctrlKind != TGB_CTRL_NONE ? w->d_ctrl : nullptr; a.status = w->d_status;
    a.sine_state = w->ctrlKind == TGB_CTRL_ESCAPE_SINE ? w->d_sine : nullptr;
    a.max_tension = w->d_max_tension;
    a.model_t = w->d_model_t; a.model_i = w->d_model_i;
    a.n_env = w->cfg.n_env; a.n_steps = nSteps;
    a.stats = w->d_stats;
    a.phases = 3;
    a.wide = wideLayout(w) ? 1 : 0;
    a.R = w->R; a.C = w->C; a.S = w->S; a.n_inc = w->nInc; a.max_inc = w->maxInc;
    a.stride = w->stride; a.cap_contacts = w->capContacts; a.frame_body = w->frameBody;
    a.ctrl_kind = w->ctrlKind; a.ctrl_stride = w->ctrlKind != TGB_CTRL_NONE ? w->ctrlStride : 0;
    a.iterations = w->cfg.solver_iterations > 0 ? w->cfg.solver_iterations : 10;
    a.ground_kind = w->ground.kind;
    a.model_t_words = w->modelTWords; a.model_i_words = w->modelIWords;
    a.model_crec_off = w->modelCrecOff; a.model_cirec_off = w->modelCirecOff;
    a.n_obst = w->nObst; a.model_obst_off = w->modelObstOff;
    a.n_pairs = w->pairCap != 0 ? w->nPairs : 0; a.model_pair_off = w->modelPairOff; a.pair_cap = w->pairCap;
    a.min_clearance = w->d_min_clearance;
    a.coupled_steps = (a.n_pairs > 0 && a.pair_cap > 0) ? w->d_coupled : nullptr;
    a.warm = w->warmFactor > 0.0 ? w->d_warm : nullptr; a.warm_factor = w->warmFactor;
    a.n_dyn = 0;
    for (int b = 0; b < w->R; ++b) a.n_dyn += w->bodies[b].mass != 0.0;
    a.dt = dt;
    a.gn[0] = w->ground.normal[0]; a.gn[1] = w->ground.normal[1]; a.gn[2] = w->ground.normal[2];
    a.goff = w->ground.offset;
    a.ground_restitution = w->ground.restitution;
    a.height_lo = w->heightLo; a.height_hi = w->heightHi;
    if (w->ctrlKind == TGB_CTRL_CPG_IMPEDANCE) {
        a.cpg_n_couplings = w->cpg.n_couplings; a.cpg_n_pairs = w->cpg.n_pairs; a.cpg_max_steps = w->cpg.max_steps;
        a.cpg_score_body = w->cpg.score_body;
        a.cpg_topo = w->d_cpg_topo; a.cpg_imp = w->d_cpg_imp; a.cpg_state = w->d_cpg_state;
        a.cpg_cable_control = w->cpg.cable_control; a.cpg_pid[0] = w->cpg.pid_kp; a.cpg_pid[1] = w->cpg.pid_ki; a.cpg_pid[2] = w->cpg.pid_kd;
        a.cpg_r_const = w->cpg.r_const; a.cpg_d_min = w->cpg.d_min; a.cpg_d_max = w->cpg.d_max; a.cpg_desc = w->cpg.descending_command;
        a.cpg_control_time = w->cpg.control_time; a.cpg_cable_step = w->cpg.cable_control_step;
        a.cpg_height_lo = w->cpg.height_lo; a.cpg_height_hi = w->cpg.height_hi;
    }
    return a;
}

/* Tail split.  A launch ends with its slowest warp, and environments whose rods lean on each other take the coupled contact solver
   (tgb_pairs.cuh) on every step for thousands of steps: a warp that holds one of them runs ~1.6x as long as the others (measured on
   the bench workload: SMs active 62 % of the launch).  Such environments are few and they stay what they are from one launch to the
   next, so a launch of many steps is cut in two by the count the last launch left per environment (tgb_step_args.coupled_steps):
   the environments that were coupled for at least 1/8 of its steps (at most splitSlowCap of them -- all CTAs of both kernels have to
   stay resident) get one warp each (32 lanes per environment: the cables run in one round instead of four, the body pairs are
   tested in one, nothing of another environment's contact code is executed), on a second stream of higher priority; everything
   else runs in the usual layout.  Both kernels take their environments from lists a small kernel writes (tgb_step_args.env_list).
   Results do not depend on which lanes simulate an environment (tested bit for bit), so the split changes durations only. */
__global__ void tgb_classify_envs_kernel(const int* coupled, int envBegin, int envCount, int threshold, int recentThreshold, int slowCap, int* count,
                                         int* slow, int* fast, int mediumThreshold, int mediumCap, int* medium) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= envCount) return;
    const int env = envBegin + i;
    const int total = coupled[env] & 0xfffff, recent = coupled[env] >> 20;
    bool heavy = false;
    if (total >= threshold || recent >= recentThreshold) {
        const int k = atomicAdd(count, 1);
        if (k < slowCap) { slow[k] = env; return; }
        heavy = true;
    }
    /* the middle class: heavy for a smaller part of the last launch (or one too many for the first list): two to a warp */
    if (mediumCap > 0 && (heavy || total >= mediumThreshold)) {
        const int k = atomicAdd(count + 2, 1);
        if (k < mediumCap) { medium[k] = env; return; }
    }
    fast[atomicAdd(count + 1, 1)] = env;
}

bool splitWanted(const tgb_world* w, const tgb_step_args& a, int envCount) {
    if (w->d_coupled == nullptr || a.coupled_steps == nullptr || a.phases != 3 || a.wide || w->lanes >= 32) return false;
    if (w->ctrlKind == TGB_CTRL_CPG_IMPEDANCE) return false; /* the CPG class gives an environment as many lanes as it has nodes */
    if (w->splitMode == 0) return false;
    if (w->splitMode == 1) return true;
    static const char* envSplit = std::getenv("TGB_TAIL_SPLIT");
    if (envSplit && std::atoi(envSplit) == 0) return false;
    /* one more kernel and two lists pay when a launch is long and the device is full enough for its slowest warp to matter */
    return a.n_steps >= 64 && envCount >= 512 && w->coupledLaunchSteps > 0;
}

tgb_status launchSplit(tgb_world* w, tgb_step_args a, int envBegin, int envCount, cudaStream_t s, int slot, int lastSteps) {
    if (!w->splitStream[slot]) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        TGB_CUDA(cudaStreamCreateWithPriority(&w->splitStream[slot], cudaStreamNonBlocking, hi));
        TGB_CUDA(cudaEventCreateWithFlags(&w->splitFork[slot], cudaEventDisableTiming));
        TGB_CUDA(cudaEventCreateWithFlags(&w->splitJoin[slot], cudaEventDisableTiming));
    }
    int* count = w->d_split_count + 4 * slot;
    int* slow = w->d_split_slow + (size_t)slot * w->splitSlowCap;
    int* medium = w->d_split_medium + (size_t)slot * w->splitMediumCap;
    /* the middle class needs 16-lane groups to be narrower than a warp's worth of the usual layout and room for its extra warps */
    /* Measured on B200 (bench workload, same box, TGB_SPLIT_MEDIUM_CAP / _DIV): off 6.57 ms per 1000 steps; cap 32 / 64 / 128 / 192 / 296
       give 7.27 / 6.99 / 7.75 / 8.48 / 8.98 -- a warp of two 16-lane environments takes longer than the usual layout's slowest warp (the
       usual-layout kernel beside it 7.8 - 8.7 ms instead of 6.5), so the class stays off; the forced mode of the tests keeps it alive. */
    const int mediumCap = w->splitMode == 1 ? std::min(w->splitMediumCap, std::max(16, splitTuning().mediumCap))
                                            : ((w->lanes <= 8 && envCount >= 1024) ? std::min(w->splitMediumCap, splitTuning().mediumCap) : 0);
    const int mediumThreshold = w->splitMode == 1 ? 1 : std::max(1, lastSteps / splitTuning().mediumDiv);
    if (mediumCap > 0 && !w->splitStream2[slot]) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        TGB_CUDA(cudaStreamCreateWithPriority(&w->splitStream2[slot], cudaStreamNonBlocking, hi));
        TGB_CUDA(cudaEventCreateWithFlags(&w->splitJoin2[slot], cudaEventDisableTiming));
    }
    int* fast = w->d_split_fast + envBegin;
    const int slowCap = w->splitMode == 1 ? std::min(w->splitSlowCap, envCount) : std::min(w->splitSlowCap, std::max(1, envCount / 20));
    const int threshold = w->splitMode == 1 ? 1 : std::max(1, lastSteps / splitTuning().totalDiv);
    /* ... or on its last steps in a row: an environment whose rods met late in the last launch starts this one so (TGB_STATS, round 2:
       launch after launch one warp of the usual layout held an environment that was coupled for 840 - 1000 of the 1000 steps and had
       been for fewer than 125 of the launch before) */
    const int recentThreshold = w->splitMode == 1 ? 1 : std::max(1, std::min(lastSteps, splitTuning().recentThreshold));
    TGB_CUDA(cudaMemsetAsync(count, 0, 4 * sizeof(int), s));
    tgb_classify_envs_kernel<<<(envCount + 255) / 256, 256, 0, s>>>(w->d_coupled, envBegin, envCount, threshold, recentThreshold, slowCap, count, slow, fast,
                                                                    mediumThreshold, mediumCap, medium);
    TGB_CUDA(cudaGetLastError());
    {
        static const char* dbg = std::getenv("TGB_SPLIT_DEBUG"); /* tuning aid: synchronises */
        if (dbg) {
            int h[4] = {0, 0, 0, 0};
            cudaMemcpyAsync(h, count, sizeof(h), cudaMemcpyDeviceToHost, s);
            cudaStreamSynchronize(s);
            std::fprintf(stderr, "tgb tail split: %d environments coupled for >= %d of the last %d steps (cap %d), %d two to a warp (>= %d, cap %d), %d in the usual layout\n", h[0],
                         threshold, lastSteps, slowCap, h[2], mediumThreshold, mediumCap, h[1]);
        }
    }
    TGB_CUDA(cudaEventRecord(w->splitFork[slot], s));
    TGB_CUDA(cudaStreamWaitEvent(w->splitStream[slot], w->splitFork[slot], 0));
    static const char* dbgT = std::getenv("TGB_SPLIT_DEBUG");
    cudaEvent_t dbgEv[4] = {nullptr, nullptr, nullptr, nullptr};
    if (dbgT) { for (int i = 0; i < 4; ++i) cudaEventCreate(&dbgEv[i]); cudaEventRecord(dbgEv[0], w->splitStream[slot]); }
    tgb_step_args as = a;
    as.env_list = slow; as.env_list_count = count; as.env_list_slots = slowCap;
    cudaError_t e = w->ops->step(as, 32, 32, w->splitStream[slot]);
    if (e != cudaSuccess) return cudaFail(e, "step kernel launch (coupled environments)");
    if (dbgT) cudaEventRecord(dbgEv[1], w->splitStream[slot]);
    TGB_CUDA(cudaEventRecord(w->splitJoin[slot], w->splitStream[slot]));
    if (mediumCap > 0) {
        TGB_CUDA(cudaStreamWaitEvent(w->splitStream2[slot], w->splitFork[slot], 0));
        tgb_step_args am = a;
        am.env_list = medium; am.env_list_count = count + 2; am.env_list_slots = mediumCap;
        e = w->ops->step(am, 16, 32, w->splitStream2[slot]);
        if (e != cudaSuccess) return cudaFail(e, "step kernel launch (two environments to a warp)");
        TGB_CUDA(cudaEventRecord(w->splitJoin2[slot], w->splitStream2[slot]));
        w->launches += 1;
    }
    tgb_step_args af = a;
    af.env_list = fast; af.env_list_count = count + 1; af.env_list_slots = envCount;
    if (dbgT) cudaEventRecord(dbgEv[2], s);
    e = w->ops->step(af, w->lanes, w->threads, s);
    if (e != cudaSuccess) return cudaFail(e, "step kernel launch");
    if (dbgT) cudaEventRecord(dbgEv[3], s);
    TGB_CUDA(cudaStreamWaitEvent(s, w->splitJoin[slot], 0));
    if (mediumCap > 0) TGB_CUDA(cudaStreamWaitEvent(s, w->splitJoin2[slot], 0));
    if (dbgT) {
        cudaStreamSynchronize(s);
        float ts = 0, tf = 0;
        cudaEventElapsedTime(&ts, dbgEv[0], dbgEv[1]); cudaEventElapsedTime(&tf, dbgEv[2], dbgEv[3]);
        std::fprintf(stderr, "tgb tail split: one warp per environment %.3f ms, usual layout %.3f ms\n", ts, tf);
        for (int i = 0; i < 4; ++i) cudaEventDestroy(dbgEv[i]);
    }
    w->launches += 3;
    return TGB_OK;
}

/* Steps per launch while the tail split is on: which environments are coupled is only known from the launch before, so a call of
   many steps is cut into launches of this many -- an environment whose rods come to rest on each other spends at most one of them in
   a warp that it holds up.  (Results do not depend on how an episode is cut into launches: tested.) */
int splitChunk(const tgb_world* w) {
    static const char* envChunk = std::getenv("TGB_SPLIT_CHUNK");
    if (envChunk && std::atoi(envChunk) > 0) return std::atoi(envChunk);
    return w->splitMode == 1 ? 100 : (1 << 30); /* measured on B200 (bench workload, 1000 steps): launches of 125 / 250 / 500 / 1000 steps 8.0 / 6.9 / 6.4 / 6.4 ms */
}

/* Worlds with contact cables: the cable's step (tgBulletContactSpringCable::step) follows the models half of every single step, so a
   call of n steps is n launches of the step kernel, each followed by one of the contact-cable kernel (tgb_ccable.cuh). */
tgb_status launchContactCableSteps(tgb_world* w, const tgb_step_args& a, const tgb::tgb_ccable_args& ca, int nSteps, cudaStream_t s) {
    for (int i = 0; i < nSteps; ++i) {
        cudaError_t e = w->ops->step(a, w->lanes, w->threads, s);
        if (e != cudaSuccess) return cudaFail(e, "step kernel launch");
        if (a.phases & 2) {
            e = w->ops->ccable_step(ca, s);
            if (e != cudaSuccess) return cudaFail(e, "contact cable kernel launch");
        }
    }
    return TGB_OK;
}

tgb_status launchWithContactCables(tgb_world* w, double dt, int nSteps, int phases, int envBegin, int envCount, cudaStream_t s) {
    tgb_step_args a = stepArgs(w, dt, 1);
    a.phases = phases;
    a.env_begin = envBegin;
    a.n_env = envBegin + envCount;
    a.coupled_steps = nullptr; /* no tail split here: launches of one step */
    const tgb::tgb_ccable_args ca = w->ccArgs(envBegin, envCount, dt);
    const int perStep = (phases & 2) ? 2 : 1;
    static const bool useGraphs = !(std::getenv("TGB_CC_GRAPHS") && std::atoi(std::getenv("TGB_CC_GRAPHS")) == 0);
    const int kChunk = 250; /* steps per graph: 500 kernel nodes; longer calls replay it */
    int done = 0;
    while (useGraphs && nSteps - done >= 8) {
        const int steps = std::min(kChunk, nSteps - done);
        cudaGraphExec_t exec = nullptr;
        for (const tgb_world::CcGraph& g : w->ccGraphs)
            if (g.steps == steps && g.envBegin == envBegin && g.envCount == envCount && std::memcmp(&g.args, &a, sizeof(a)) == 0) { exec = g.exec; break; }
        if (!exec) {
            /* one untimed direct step first: the kernels' function attributes are set outside the capture */
            if (!w->ccLaunchedDirectly) {
                tgb_status st = launchContactCableSteps(w, a, ca, 1, s);
                if (st != TGB_OK) return st;
                w->ccLaunchedDirectly = true;
                w->launches += perStep;
                ++done;
                continue;
            }
            if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); break; }
            const tgb_status st = launchContactCableSteps(w, a, ca, steps, s);
            cudaGraph_t graph = nullptr;
            const cudaError_t ee = cudaStreamEndCapture(s, &graph);
            if (st != TGB_OK || ee != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); break; }
            const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ei != cudaSuccess) { cudaGetLastError(); break; }
            if (w->ccGraphs.size() >= 16) { /* the oldest graph may still be replaying on another stream of this world */
                cudaDeviceSynchronize();
                cudaGraphExecDestroy(w->ccGraphs.front().exec);
                w->ccGraphs.erase(w->ccGraphs.begin());
            }
            tgb_world::CcGraph g;
            g.args = a; g.steps = steps; g.envBegin = envBegin; g.envCount = envCount; g.exec = exec;
            w->ccGraphs.push_back(g);
        }
        const cudaError_t e = cudaGraphLaunch(exec, s);
        if (e != cudaSuccess) return cudaFail(e, "contact cable graph launch");
        w->launches += (int64_t)perStep * steps;
        done += steps;
    }
    if (done < nSteps) {
        tgb_status st = launchContactCableSteps(w, a, ca, nSteps - done, s);
        if (st != TGB_OK) return st;
        w->launches += (int64_t)perStep * (nSteps - done);
    }
    return TGB_OK;
}

tgb_status launchStep(tgb_world* w, double dt, int nSteps, int phases = 3) {
    if (!w->contactCables.empty()) {
        tgb_status st = launchWithContactCables(w, dt, nSteps, phases, 0, w->cfg.n_env, w->stream);
        if (st == TGB_OK && (phases & 1)) w->time += dt * nSteps;
        return st;
    }
    if (w->d_stats && std::getenv("TGB_STATS_RESET")) { /* instrumentation: the counters of the last launch only */
        cudaMemsetAsync(w->d_stats, 0, (128 + 8 * 4096) * sizeof(unsigned long long), w->stream);
        cudaMemsetAsync(w->d_stats + 53, 0xff, sizeof(unsigned long long), w->stream);
    }
    tgb_step_args a = stepArgs(w, dt, nSteps);
    a.phases = phases;
    if (splitWanted(w, a, w->cfg.n_env)) {
        const int chunk = splitChunk(w);
        for (int done = 0; done < nSteps; done += chunk) {
            a.n_steps = std::min(chunk, nSteps - done);
            tgb_status st = launchSplit(w, a, 0, w->cfg.n_env, w->stream, 0, w->coupledLaunchSteps);
            if (st != TGB_OK) return st;
            w->coupledLaunchSteps = a.n_steps;
        }
    } else {
        cudaError_t e = w->ops->step(a, w->lanes, w->threads, w->stream);
        if (e != cudaSuccess) return cudaFail(e, "step kernel launch");
        w->launches += 1;
        if ((phases & 1) && a.coupled_steps) w->coupledLaunchSteps = nSteps;
    }
    if (phases & 1) w->time += dt * nSteps;
    return TGB_OK;
}

/* the same for the environments [envBegin, envBegin + envCount) only, on stream s; does not advance the world's clock.  slot: which
   of the world's list buffers the range may use (1 + chunk of tgb_rollout_host) */
tgb_status launchStepRange(tgb_world* w, double dt, int nSteps, int envBegin, int envCount, cudaStream_t s, int slot) {
    if (!w->contactCables.empty()) return launchWithContactCables(w, dt, nSteps, 3, envBegin, envCount, s);
    tgb_step_args a = stepArgs(w, dt, nSteps);
    if (splitWanted(w, a, envCount)) {
        const int chunk = splitChunk(w);
        int lastSteps = w->coupledLaunchSteps;
        for (int done = 0; done < nSteps; done += chunk) {
            a.n_steps = std::min(chunk, nSteps - done);
            tgb_status st = launchSplit(w, a, envBegin, envCount, s, slot, lastSteps);
            if (st != TGB_OK) return st;
            lastSteps = a.n_steps;
        }
        w->coupledLaunchSteps = std::min(chunk, nSteps); /* (the chunks of one call of tgb_rollout_host are cut alike) */
    } else {
        a.env_begin = envBegin;
        a.n_env = envBegin + envCount;
        cudaError_t e = w->ops->step(a, w->lanes, w->threads, s);
        if (e != cudaSuccess) return cudaFail(e, "step kernel launch");
        w->launches += 1;
        if (a.coupled_steps) w->coupledLaunchSteps = nSteps;
    }
    return TGB_OK;
}

} // namespace

extern "C" {

int tgb_abi_version(void) { return TGB_ABI_VERSION; }
const char* tgb_last_error(void) { return g_error.c_str(); }
int tgb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

tgb_status tgb_create(const tgb_world_cfg* cfg, const tgb_ground_desc* ground, tgb_world** out) {
    if (!cfg || !ground || !out) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_create: NULL argument");
    *out = nullptr;
    if (cfg->n_env < 1) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_create: n_env must be positive");
    if (cfg->precision != TGB_F32 && cfg->precision != TGB_F64) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_create: unknown precision");
    if (ground->kind < TGB_GROUND_EMPTY || ground->kind > TGB_GROUND_PLANE) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_create: unknown ground kind");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(TGB_ERR_CUDA, "tgb_create: no CUDA device is available (this library has no CPU path)");
    }
    if (cfg->device < 0 || cfg->device >= n) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_create: device ordinal out of range");
    TGB_CUDA(cudaSetDevice(cfg->device));
    tgb_world* w = new tgb_world();
    w->cfg = *cfg;
    w->ground = *ground;
    w->ops = cfg->precision == TGB_F64 ? tgb_ops_f64() : tgb_ops_f32();
    w->scalarBytes = w->ops->scalar_bytes;
    { int sms = 0; if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && sms > 0) w->smCount = sms; }
    e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&w->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&w->ev1);
    if (e != cudaSuccess) { w->release(); delete w; return cudaFail(e, "tgb_create: stream/event creation"); }
    if (std::getenv("TGB_STATS")) {
        /* 128 global counters, then 8 per CTA (first warp) for up to 4096 CTAs: see the TGB_STATS blocks of tgb_step.cuh */
        if (cudaMalloc(&w->d_stats, (128 + 8 * 4096) * sizeof(unsigned long long)) == cudaSuccess) {
            cudaMemset(w->d_stats, 0, (128 + 8 * 4096) * sizeof(unsigned long long));
            cudaMemset(w->d_stats + 53, 0xff, sizeof(unsigned long long));
        }
    }
    *out = w;
    return TGB_OK;
}

tgb_status tgb_destroy(tgb_world* w) {
    if (!w) return TGB_OK;
    cudaSetDevice(w->cfg.device);
    if (w->stream) cudaStreamSynchronize(w->stream);
    if (w->d_stats) {
        std::vector<unsigned long long> h(128 + 8 * 4096);
        if (cudaMemcpy(h.data(), w->d_stats, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
            std::fprintf(stderr, "TGB_STATS");
            for (int i = 0; i < 128; ++i) std::fprintf(stderr, " %llu", h[i]);
            std::fprintf(stderr, "\n");
            /* per CTA: total cycles, coupled warp-steps, cycles in the coupled solve, cycles in the pair detection, warp-steps with the
               shared-memory solver, warp-steps with a lane of three or more contact points, cycles in the cable phase, spare */
            for (int c = 0; c < 4096; ++c) {
                const unsigned long long* r = h.data() + 128 + 8 * c;
                if (r[0] == 0) continue;
                std::fprintf(stderr, "TGB_CTA %d %llu %llu %llu %llu %llu %llu %llu %llu\n", c, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
            }
        }
        cudaFree(w->d_stats);
    }
    w->release();
    delete w;
    return TGB_OK;
}

tgb_status tgb_add_model_flat(tgb_world* w, const tgb_body_desc* bodies, int32_t n_bodies, const tgb_shape_desc* shapes, int32_t n_shapes,
                              const tgb_cable_desc* cables, int32_t n_cables) {
    if (!w) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_add_model_flat: NULL world");
    if (w->finalized) return fail(TGB_ERR_STATE, "tgb_add_model_flat: world already finalized");
    if (n_bodies < 0 || n_shapes < 0 || n_cables < 0 || (n_bodies && !bodies) || (n_shapes && !shapes) || (n_cables && !cables))
        return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_add_model_flat: bad array arguments");
    const int body0 = (int)w->bodies.size(), shape0 = (int)w->shapes.size();
    w->lastModelBody0 = body0;
    for (int i = 0; i < n_bodies; ++i) {
        const tgb_body_desc& b = bodies[i];
        if (b.mass < 0.0 || b.shape_begin < 0 || b.shape_count < 0 || b.shape_begin + b.shape_count > n_shapes)
            return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_add_model_flat: bad body descriptor");
    }
    for (int i = 0; i < n_cables; ++i) {
        const tgb_cable_desc& c = cables[i];
        if (c.body_a < 0 || c.body_a >= n_bodies || c.body_b < 0 || c.body_b >= n_bodies)
            return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_add_model_flat: cable body index out of range");
        /* same checks as tgSpringCableActuator::Config (tgSpringCableActuator.cpp:61-87) */
        if (c.stiffness < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "stiffness is negative.");
        if (c.damping < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "damping is negative.");
        if (c.max_tension < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "max tension is negative.");
        if (c.min_actual_length < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "min Actual Length is negative.");
        if (c.min_rest_length < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "min Rest Length is negative.");
        if (c.actuator < TGB_ACT_PASSIVE || c.actuator > TGB_ACT_KINEMATIC) return fail(TGB_ERR_INVALID_ARGUMENT, "unknown actuator kind");
        if (!(c.rest_length > 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "Pretension causes string to shorten past rest length!");
    }
    for (int i = 0; i < n_shapes; ++i) {
        if (shapes[i].kind < TGB_SHAPE_CYLINDER || shapes[i].kind > TGB_SHAPE_BOX) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_add_model_flat: unknown shape kind");
        if (shapes[i].radius < 0.0 || shapes[i].half_length < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "Negative radius");
    }
    for (int i = 0; i < n_shapes; ++i) w->shapes.push_back(shapes[i]);
    for (int i = 0; i < n_bodies; ++i) {
        tgb_body_desc b = bodies[i];
        b.shape_begin += shape0;
        w->bodies.push_back(b);
    }
    for (int i = 0; i < n_cables; ++i) {
        tgb_cable_desc c = cables[i];
        c.body_a += body0; c.body_b += body0;
        w->cables.push_back(c);
    }
    return TGB_OK;
}

tgb_status tgb_add_contact_cables(tgb_world* w, const tgb_contact_cable_desc* cables, int32_t n_cables) {
    if (!w) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_add_contact_cables: NULL world");
    if (w->finalized) return fail(TGB_ERR_STATE, "tgb_add_contact_cables: world already finalized");
    if (n_cables < 0 || (n_cables && !cables)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_add_contact_cables: bad array arguments");
    const int nb = (int)w->bodies.size() - w->lastModelBody0;
    for (int i = 0; i < n_cables; ++i) {
        const tgb_contact_cable_desc& c = cables[i];
        if (c.body_a < 0 || c.body_a >= nb || c.body_b < 0 || c.body_b >= nb)
            return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_add_contact_cables: cable body index out of range");
        /* tgBulletContactSpringCable::invariant (tgBulletContactSpringCable.cpp:939-952) */
        if (!(c.stiffness > 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "stiffness is not positive.");
        if (c.damping < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "damping is negative.");
        if (c.rest_length < 0.0 || c.prev_length < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "Pretension causes string to shorten past rest length!");
        if (c.thickness < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "thickness is negative.");
        if (!(c.resolution > 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "resolution is not positive.");
        /* tgSpringCableActuator::Config (tgSpringCableActuator.cpp:61-87) */
        if (c.max_tension < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "max tension is negative.");
        if (c.min_actual_length < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "min Actual Length is negative.");
        if (c.min_rest_length < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "min Rest Length is negative.");
    }
    for (int i = 0; i < n_cables; ++i) {
        tgb_contact_cable_desc c = cables[i];
        c.body_a += w->lastModelBody0; c.body_b += w->lastModelBody0;
        w->contactCables.push_back(c);
    }
    return TGB_OK;
}
tgb_status tgb_num_contact_cables(const tgb_world* w, int32_t* out) {
    if (!w || !out) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = (int32_t)w->contactCables.size();
    return TGB_OK;
}

tgb_status tgb_set_lanes_per_env(tgb_world* w, int32_t lanes) {
    if (!w) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL world");
    if (w->finalized) return fail(TGB_ERR_STATE, "tgb_set_lanes_per_env: world already finalized");
    if (lanes != 0 && lanes != 2 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32)
        return fail(TGB_ERR_INVALID_ARGUMENT, "lanes per env must be 0 (automatic), 2, 4, 8, 16 or 32");
    w->lanesOverride = lanes;
    return TGB_OK;
}

tgb_status tgb_set_wide_layout(tgb_world* w, int32_t mode) {
    if (!w) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL world");
    if (mode < -1 || mode > 1) return fail(TGB_ERR_INVALID_ARGUMENT, "layout mode must be -1 (automatic), 0 (narrow) or 1 (wide)");
    w->wideMode = mode;
    return TGB_OK;
}

tgb_status tgb_set_tail_split(tgb_world* w, int32_t mode) {
    if (!w) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL world");
    if (mode < -1 || mode > 1) return fail(TGB_ERR_INVALID_ARGUMENT, "tail split mode must be -1 (automatic), 0 (off) or 1 (every launch, any count)");
    w->splitMode = mode;
    return TGB_OK;
}

tgb_status tgb_set_pair_contact_cap(tgb_world* w, int32_t cap) {
    if (!w) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL world");
    if (w->finalized) return fail(TGB_ERR_STATE, "tgb_set_pair_contact_cap: world already finalized");
    if (cap < -1 || cap > 64) return fail(TGB_ERR_INVALID_ARGUMENT, "pair contact cap must be -1 (measure the clearance only), 0 (off) .. 64");
    w->pairCap = cap;
    return TGB_OK;
}

tgb_status tgb_set_warm_starting(tgb_world* w, double factor) {
    if (!w) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL world");
    if (!(factor >= 0.0 && factor <= 1.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "warm starting factor must be in [0, 1]");
    if (w->finalized && factor > 0.0) {
        /* the persistent manifolds take shared memory in the step kernel: refuse rather than fail at the next launch */
        const double before = w->warmFactor;
        w->warmFactor = factor;
        tgb_step_args a = stepArgs(w, 1e-3, 1);
        if (a.ctrl_stride == 0) a.ctrl_stride = std::max(w->C, TGB_ESCAPE_PARAMS);
        if (w->ops->step_smem(a, w->lanes, w->threads) > 227 * 1024) {
            w->warmFactor = before;
            return fail(TGB_ERR_UNSUPPORTED, "tgb_set_warm_starting: the persistent manifolds do not fit in shared memory for this model");
        }
    }
    w->warmFactor = factor;
    if (w->finalized && w->d_warm) {
        /* a new factor starts from empty manifolds */
        TGB_CUDA(cudaSetDevice(w->cfg.device));
        TGB_CUDA(cudaStreamSynchronize(w->stream));
        const size_t words = (size_t)w->cfg.n_env * w->warmWords;
        if (w->scalarBytes == 4) {
            std::vector<float> empty(words, -1.0f);
            TGB_CUDA(cudaMemcpy(w->d_warm, empty.data(), words * sizeof(float), cudaMemcpyHostToDevice));
        } else {
            std::vector<double> empty(words, -1.0);
            TGB_CUDA(cudaMemcpy(w->d_warm, empty.data(), words * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    return TGB_OK;
}

tgb_status tgb_get_min_clearance(tgb_world* w, int32_t env_begin, int32_t env_count, double* out, int32_t reset) {
    if (!w) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL world");
    if (!w->finalized) return fail(TGB_ERR_STATE, "world is not finalized");
    TGB_CUDA(cudaSetDevice(w->cfg.device));
    if (!out || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_get_min_clearance: bad range or NULL buffer");
    if (env_count == 0) return TGB_OK;
    std::vector<float> tmp((size_t)env_count);
    TGB_CUDA(cudaMemcpyAsync(tmp.data(), w->d_min_clearance + env_begin, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    for (int i = 0; i < env_count; ++i) out[i] = (double)tmp[i];
    if (reset) {
        for (int i = 0; i < env_count; ++i) tmp[i] = 1e30f;
        TGB_CUDA(cudaMemcpyAsync(w->d_min_clearance + env_begin, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice, w->stream));
        TGB_CUDA(cudaStreamSynchronize(w->stream));
    }
    return TGB_OK;
}

tgb_status tgb_set_block_threads(tgb_world* w, int32_t threads) {
    if (!w) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL world");
    if (threads != 0 && threads != 32 && threads != 64 && threads != 96 && threads != 128)
        return fail(TGB_ERR_INVALID_ARGUMENT, "block threads must be 0 (automatic), 32, 64, 96 or 128");
    w->threadsOverride = threads;
    if (w->finalized && threads) {
        if (threads % w->lanes) return fail(TGB_ERR_INVALID_ARGUMENT, "block threads must be a multiple of lanes per env");
        w->threads = threads;
    }
    return TGB_OK;
}

tgb_status tgb_finalize(tgb_world* w) {
    if (!w) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL world");
    if (w->finalized) return TGB_OK;
    w->R = (int)w->bodies.size(); w->C = (int)w->cables.size(); w->S = (int)w->shapes.size();
    if (w->R < 1) return fail(TGB_ERR_STATE, "tgb_finalize: the world has no bodies");
    if (w->R > 32) return fail(TGB_ERR_UNSUPPORTED, "tgb_finalize: more than 32 rigid bodies per environment");
    if (w->C >= (1 << 24)) return fail(TGB_ERR_UNSUPPORTED, "tgb_finalize: too many cables");
    TGB_CUDA(cudaSetDevice(w->cfg.device));
    const int n = w->cfg.n_env;
    const int vecWords = (int)(16 / w->scalarBytes);
    w->stride = TGB_CABLE_BASE(w->R) + w->C * TGB_NCS;
    w->stride = (w->stride + vecWords - 1) / vecWords * vecWords;
    int maxShapes = 1;
    /* the env frame follows the first static body if there is one (then it never moves and static bodies stay
       bit-exactly in place), else the first body */
    w->frameBody = 0;
    for (int b = w->R - 1; b >= 0; --b) {
        if (w->bodies[b].mass == 0.0) w->frameBody = b;
        else maxShapes = std::max(maxShapes, (int)w->bodies[b].shape_count); /* contact rows exist for dynamic bodies only */
    }
    w->capContacts = 4 * maxShapes; /* at most 4 manifold points per collision shape */
    w->lanes = w->lanesOverride ? w->lanesOverride : std::max(2, nextPow2(w->R));
    if (w->lanes < w->R) return fail(TGB_ERR_INVALID_ARGUMENT, "lanes per env is smaller than the number of bodies");
    /* block size: enough CTAs to spread over all SMs; a group never spans warps */
    int threads = w->threadsOverride;
    if (!threads) {
        const char* envThreads = std::getenv("TGB_BLOCK_THREADS");
        if (envThreads) threads = std::atoi(envThreads);
    }
    if (!threads) {
        const long long totalThreads = (long long)n * w->lanes;
        threads = totalThreads >= 148LL * 16 * 128 ? 128 : totalThreads >= 148LL * 16 * 64 ? 64 : 32;
    }
    if (threads % 32 || threads < 32 || threads > 128) return fail(TGB_ERR_INVALID_ARGUMENT, "bad block size");
    w->threads = threads;

    {
        /* tuning aid: TGB_PAIR_CAP overrides the pair contact cap (0 = off, -1 = detection only: pairs are tested, no row is emitted) */
        const char* envCap = std::getenv("TGB_PAIR_CAP");
        if (envCap) w->pairCap = std::max(-1, std::min(64, std::atoi(envCap)));
    }
    std::vector<unsigned char> tbytes;
    std::vector<int> ints;
    if (w->cfg.precision == TGB_F64) packModel<double>(w, tbytes, ints, w->nInc, w->modelCrecOff, w->modelCirecOff, w->maxInc, w->nObst, w->modelObstOff, w->nPairs, w->modelPairOff);
    else packModel<float>(w, tbytes, ints, w->nInc, w->modelCrecOff, w->modelCirecOff, w->maxInc, w->nObst, w->modelObstOff, w->nPairs, w->modelPairOff);
    /* obstacle contacts share the rows: room for those of up to two obstacles touched at once */
    w->capContacts *= 1 + std::min(w->nObst, 2);
    w->modelTWords = (int)(tbytes.size() / w->scalarBytes);
    w->modelIWords = (int)ints.size();
    TGB_CUDA(cudaMalloc(&w->d_model_t, std::max<size_t>(tbytes.size(), 16)));
    TGB_CUDA(cudaMalloc(&w->d_model_i, std::max<size_t>(ints.size() * sizeof(int), 16)));
    TGB_CUDA(cudaMemcpy(w->d_model_t, tbytes.data(), tbytes.size(), cudaMemcpyHostToDevice));
    TGB_CUDA(cudaMemcpy(w->d_model_i, ints.data(), ints.size() * sizeof(int), cudaMemcpyHostToDevice));
    TGB_CUDA(cudaMalloc(&w->d_state, (size_t)n * w->stride * w->scalarBytes));
    TGB_CUDA(cudaMemset(w->d_state, 0, (size_t)n * w->stride * w->scalarBytes));
    TGB_CUDA(cudaMalloc(&w->d_frame, (size_t)n * TGB_NFR * sizeof(double)));
    TGB_CUDA(cudaMemset(w->d_frame, 0, (size_t)n * TGB_NFR * sizeof(double)));
    TGB_CUDA(cudaMalloc(&w->d_com0, (size_t)n * 3 * sizeof(double)));
    TGB_CUDA(cudaMalloc(&w->d_status, (size_t)n * sizeof(int)));
    TGB_CUDA(cudaMemset(w->d_status, 0, (size_t)n * sizeof(int)));
    {
        std::vector<float> far((size_t)n, 1e30f);
        TGB_CUDA(cudaMalloc(&w->d_min_clearance, (size_t)n * sizeof(float)));
        TGB_CUDA(cudaMemcpy(w->d_min_clearance, far.data(), (size_t)n * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (w->nPairs > 0 && w->pairCap > 0) {
        /* tail split: all CTAs of both kernels of a launch have to be resident (8 one-warp CTAs per SM), so the environments that get a
           warp of their own are capped at what the device has left over */
        /* (measured: with the cap of 148 reached the launch took twice as long -- not every CTA of the second kernel found a place
           beside seven or eight of the first; half of that has always fitted) */
        w->splitSlowCap = std::max(16, std::min(n, w->smCount / 2));
        TGB_CUDA(cudaMalloc(&w->d_coupled, (size_t)n * sizeof(int)));
        TGB_CUDA(cudaMemset(w->d_coupled, 0, (size_t)n * sizeof(int)));
        w->splitMediumCap = std::max(16, std::min(n, 2 * w->smCount));
        TGB_CUDA(cudaMalloc(&w->d_split_medium, (size_t)tgb_world::kSplitSlots * w->splitMediumCap * sizeof(int)));
        TGB_CUDA(cudaMalloc(&w->d_split_count, 4 * tgb_world::kSplitSlots * sizeof(int)));
        TGB_CUDA(cudaMalloc(&w->d_split_slow, (size_t)tgb_world::kSplitSlots * w->splitSlowCap * sizeof(int)));
        TGB_CUDA(cudaMalloc(&w->d_split_fast, (size_t)n * sizeof(int)));
    }
    {
        /* persistent manifolds: empty (key -1) */
        int nDyn = 0;
        for (int b = 0; b < w->R; ++b) nDyn += w->bodies[b].mass != 0.0;
        const int rowBodies = w->nObst > 0 ? nDyn : w->R;
        const int pairCapRows = (w->pairCap > 0 && w->nPairs > 0) ? w->pairCap : 0;
        w->warmWords = TGB_WARM_STATE_WORDS(w->capContacts, rowBodies, pairCapRows);
        const size_t words = (size_t)n * w->warmWords;
        TGB_CUDA(cudaMalloc(&w->d_warm, std::max<size_t>(words * w->scalarBytes, 16)));
        if (w->scalarBytes == 4) {
            std::vector<float> empty(words, -1.0f);
            TGB_CUDA(cudaMemcpy(w->d_warm, empty.data(), words * sizeof(float), cudaMemcpyHostToDevice));
        } else {
            std::vector<double> empty(words, -1.0);
            TGB_CUDA(cudaMemcpy(w->d_warm, empty.data(), words * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    TGB_CUDA(cudaMalloc(&w->d_scores, (size_t)n * 2 * sizeof(float)));
    TGB_CUDA(cudaMemset(w->d_scores, 0, (size_t)n * 2 * sizeof(float)));

    /* the step kernel must fit its shared memory: a CTA holds threads / lanes environments, so a model with many collision
       shapes per body (contact rows) gets more lanes per environment until a CTA fits */
    {
        tgb_step_args a = stepArgs(w, 1e-3, 1);
        a.ctrl_stride = std::max(w->C, TGB_ESCAPE_PARAMS);
        while (w->ops->step_smem(a, w->lanes, w->threads) > 227 * 1024 && w->lanes < 32 && !w->lanesOverride) w->lanes *= 2;
        const size_t smem = w->ops->step_smem(a, w->lanes, w->threads);
        if (smem > 227 * 1024) return fail(TGB_ERR_UNSUPPORTED, "tgb_finalize: model too large for one CTA's shared memory");
    }

    if (!w->contactCables.empty()) {
        const int ncc = (int)w->contactCables.size();
        std::vector<double> cm((size_t)ncc * tgb::TGB_CC_MODEL, 0.0), cs((size_t)ncc * tgb::TGB_CC_STATE, 0.0);
        for (int i = 0; i < ncc; ++i) {
            const tgb_contact_cable_desc& c = w->contactCables[i];
            double* m = &cm[(size_t)i * tgb::TGB_CC_MODEL];
            m[tgb::CCM_A] = c.body_a; m[tgb::CCM_B] = c.body_b;
            for (int k = 0; k < 3; ++k) { m[tgb::CCM_LA + k] = c.local_a[k]; m[tgb::CCM_LB + k] = c.local_b[k]; }
            m[tgb::CCM_K] = c.stiffness; m[tgb::CCM_C] = c.damping; m[tgb::CCM_THICK] = c.thickness; m[tgb::CCM_RES] = c.resolution;
            m[tgb::CCM_MAXTENS] = c.max_tension; m[tgb::CCM_TVEL] = c.target_velocity; m[tgb::CCM_MINAL] = c.min_actual_length;
            m[tgb::CCM_MINRL] = c.min_rest_length;
            cs[(size_t)i * tgb::TGB_CC_STATE + tgb::CCS_PREF] = c.rest_length; /* tgBasicActuator ctor: preferred = rest */
            cs[(size_t)i * tgb::TGB_CC_STATE + tgb::CCS_REST] = c.rest_length;
            cs[(size_t)i * tgb::TGB_CC_STATE + tgb::CCS_PREV] = c.prev_length;
        }
        TGB_CUDA(cudaMalloc(&w->d_cc_model, cm.size() * sizeof(double)));
        TGB_CUDA(cudaMemcpy(w->d_cc_model, cm.data(), cm.size() * sizeof(double), cudaMemcpyHostToDevice));
        std::vector<double> all((size_t)n * cs.size());
        for (int e = 0; e < n; ++e) std::copy(cs.begin(), cs.end(), all.begin() + (size_t)e * cs.size());
        TGB_CUDA(cudaMalloc(&w->d_cc_state, all.size() * sizeof(double)));
        TGB_CUDA(cudaMemcpy(w->d_cc_state, all.data(), all.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    /* replicate the model's initial state into every environment */
    std::vector<double> rods((size_t)w->R * TGB_NBS, 0.0), cab((size_t)std::max(1, w->C) * 2, 0.0);
    for (int b = 0; b < w->R; ++b) {
        double* s = &rods[(size_t)b * TGB_NBS];
        for (int k = 0; k < 3; ++k) s[k] = w->bodies[b].pos[k];
        for (int k = 0; k < 4; ++k) s[3 + k] = w->bodies[b].quat[k];
    }
    for (int c = 0; c < w->C; ++c) { cab[2 * c] = w->cables[c].rest_length; cab[2 * c + 1] = w->cables[c].prev_length; }
    tgb_status st = w->ensureStage(rods.size() + cab.size());
    if (st != TGB_OK) return st;
    TGB_CUDA(cudaMemcpyAsync(w->d_stage, rods.data(), rods.size() * sizeof(double), cudaMemcpyHostToDevice, w->stream));
    TGB_CUDA(cudaMemcpyAsync(w->d_stage + rods.size(), cab.data(), cab.size() * sizeof(double), cudaMemcpyHostToDevice, w->stream));
    const tgb::tgb_io_args io = w->ioArgs(0, n);
    TGB_CUDA(w->ops->pack_rods(io, w->d_stage, 1, w->stream));
    TGB_CUDA(w->ops->pack_cables(io, w->d_stage + rods.size(), 1, 1, w->stream));
    TGB_CUDA(w->ops->com(io, 0, nullptr, nullptr, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    w->finalized = true;
    return TGB_OK;
}

#define TGB_NEED_FINAL(w)                                                          \
    do {                                                                           \
        if (!(w)) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL world");             \
        if (!(w)->finalized) return fail(TGB_ERR_STATE, "world is not finalized"); \
        TGB_CUDA(cudaSetDevice((w)->cfg.device));                                  \
    } while (0)

tgb_status tgb_num_bodies(const tgb_world* w, int32_t* out) {
    if (!w || !out) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = (int32_t)w->bodies.size();
    return TGB_OK;
}
tgb_status tgb_num_cables(const tgb_world* w, int32_t* out) {
    if (!w || !out) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = (int32_t)w->cables.size();
    return TGB_OK;
}
tgb_status tgb_num_envs(const tgb_world* w, int32_t* out) {
    if (!w || !out) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = w->cfg.n_env;
    return TGB_OK;
}

tgb_status tgb_set_rod_state(tgb_world* w, int32_t env_begin, int32_t env_count, const double* state) {
    TGB_NEED_FINAL(w);
    if (!state || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_rod_state: bad range or NULL buffer");
    if (env_count == 0) return TGB_OK;
    const size_t n = (size_t)env_count * w->R * TGB_NBS;
    tgb_status st = w->ensureStage(n);
    if (st != TGB_OK) return st;
    TGB_CUDA(cudaMemcpyAsync(w->d_stage, state, n * sizeof(double), cudaMemcpyHostToDevice, w->stream));
    const tgb::tgb_io_args io = w->ioArgs(env_begin, env_count);
    TGB_CUDA(w->ops->pack_rods(io, w->d_stage, 0, w->stream));
    TGB_CUDA(w->ops->com(io, 0, nullptr, nullptr, w->stream)); /* a new state starts a new episode for the scores */
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_get_rod_state(tgb_world* w, int32_t env_begin, int32_t env_count, double* state) {
    TGB_NEED_FINAL(w);
    if (!state || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_get_rod_state: bad range or NULL buffer");
    if (env_count == 0) return TGB_OK;
    const size_t n = (size_t)env_count * w->R * TGB_NBS;
    tgb_status st = w->ensureStage(n);
    if (st != TGB_OK) return st;
    TGB_CUDA(w->ops->unpack_rods(w->ioArgs(env_begin, env_count), w->d_stage, w->stream));
    TGB_CUDA(cudaMemcpyAsync(state, w->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_get_contact_cable_obs(tgb_world* w, int32_t env_begin, int32_t env_count, double* obs) {
    TGB_NEED_FINAL(w);
    if (!obs || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_get_contact_cable_obs: bad range or NULL buffer");
    const size_t n = (size_t)env_count * w->contactCables.size() * TGB_CCABLE_OBS_DOUBLES;
    if (n == 0) return TGB_OK;
    tgb_status st = w->ensureStage(n);
    if (st != TGB_OK) return st;
    TGB_CUDA(w->ops->ccable_obs(w->ccArgs(env_begin, env_count, 0.0), w->d_stage, -1, nullptr, w->stream));
    TGB_CUDA(cudaMemcpyAsync(obs, w->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_reset_contact_cables(tgb_world* w, int32_t env_begin, int32_t env_count, const double* rest_prev) {
    TGB_NEED_FINAL(w);
    if (!rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_reset_contact_cables: bad range");
    const size_t ncc = w->contactCables.size();
    if (ncc == 0 || env_count == 0) return TGB_OK;
    std::vector<double> all((size_t)env_count * ncc * tgb::TGB_CC_STATE, 0.0);
    for (int e = 0; e < env_count; ++e)
        for (size_t i = 0; i < ncc; ++i) {
            double* cs = &all[((size_t)e * ncc + i) * tgb::TGB_CC_STATE];
            const double rest = rest_prev ? rest_prev[((size_t)e * ncc + i) * 2] : w->contactCables[i].rest_length;
            const double prev = rest_prev ? rest_prev[((size_t)e * ncc + i) * 2 + 1] : w->contactCables[i].prev_length;
            if (rest < 0.0 || prev < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_reset_contact_cables: lengths must not be negative");
            cs[tgb::CCS_REST] = rest; cs[tgb::CCS_PREV] = prev; cs[tgb::CCS_PREF] = rest;
        }
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    TGB_CUDA(cudaMemcpy(w->d_cc_state + (size_t)env_begin * ncc * tgb::TGB_CC_STATE, all.data(), all.size() * sizeof(double), cudaMemcpyHostToDevice));
    return TGB_OK;
}

tgb_status tgb_set_contact_cable_control_input(tgb_world* w, int32_t env_begin, int32_t env_count, const double* input, int32_t broadcast,
                                               double dt) {
    TGB_NEED_FINAL(w);
    if (!input || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_contact_cable_control_input: bad range or NULL buffer");
    const size_t ncc = w->contactCables.size();
    const size_t n = (broadcast ? 1 : (size_t)env_count) * ncc;
    if (n == 0 || env_count == 0) return TGB_OK;
    for (size_t i = 0; i < n; ++i)
        if (input[i] < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "Rest length is negative."); /* tgBasicActuator.cpp:139-142 */
    tgb_status st = w->ensureStage(n);
    if (st != TGB_OK) return st;
    TGB_CUDA(cudaMemcpyAsync(w->d_stage, input, n * sizeof(double), cudaMemcpyHostToDevice, w->stream));
    TGB_CUDA(w->ops->ccable_control(w->ccArgs(env_begin, env_count, dt), w->d_stage, broadcast ? 1 : 0, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_get_contact_cable_anchors(tgb_world* w, int32_t cable, int32_t env_begin, int32_t env_count, double* out) {
    TGB_NEED_FINAL(w);
    if (!out || !rangeOk(w, env_begin, env_count) || cable < 0 || cable >= (int)w->contactCables.size())
        return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_get_contact_cable_anchors: bad cable, range or NULL buffer");
    const size_t n = (size_t)env_count * TGB_CCABLE_MAX_SLIDING * TGB_CCABLE_ANCHOR_DOUBLES;
    if (n == 0) return TGB_OK;
    tgb_status st = w->ensureStage(n);
    if (st != TGB_OK) return st;
    TGB_CUDA(w->ops->ccable_obs(w->ccArgs(env_begin, env_count, 0.0), nullptr, cable, w->d_stage, w->stream));
    TGB_CUDA(cudaMemcpyAsync(out, w->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_set_cable_state(tgb_world* w, int32_t env_begin, int32_t env_count, const double* rest_prev) {
    TGB_NEED_FINAL(w);
    if (!rest_prev || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_cable_state: bad range or NULL buffer");
    if (env_count == 0 || w->C == 0) return TGB_OK;
    const size_t n = (size_t)env_count * w->C * 2;
    for (size_t i = 0; i < n; i += 2)
        if (!(rest_prev[i] > 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_cable_state: rest length must be positive");
    tgb_status st = w->ensureStage(n);
    if (st != TGB_OK) return st;
    TGB_CUDA(cudaMemcpyAsync(w->d_stage, rest_prev, n * sizeof(double), cudaMemcpyHostToDevice, w->stream));
    TGB_CUDA(w->ops->pack_cables(w->ioArgs(env_begin, env_count), w->d_stage, 0, 0, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_get_cable_obs(tgb_world* w, int32_t env_begin, int32_t env_count, double* obs) {
    TGB_NEED_FINAL(w);
    if (!obs || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_get_cable_obs: bad range or NULL buffer");
    if (env_count == 0 || w->C == 0) return TGB_OK;
    const size_t n = (size_t)env_count * w->C * TGB_CABLE_OBS_DOUBLES;
    tgb_status st = w->ensureStage(n);
    if (st != TGB_OK) return st;
    TGB_CUDA(w->ops->cable_obs(w->ioArgs(env_begin, env_count), w->d_stage, w->stream));
    TGB_CUDA(cudaMemcpyAsync(obs, w->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_set_controller(tgb_world* w, int32_t kind, const double* params, int32_t params_per_env, int32_t env_begin, int32_t env_count,
                              int32_t broadcast) {
    TGB_NEED_FINAL(w);
    if (!rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_controller: bad env range");
    int need = 0;
    switch (kind) {
    case TGB_CTRL_NONE: need = 0; break;
    case TGB_CTRL_SETPOINT: case TGB_CTRL_TORQUE: need = w->C; break;
    case TGB_CTRL_TENSION: need = 1; break;
    case TGB_CTRL_ESCAPE_SINE: need = TGB_ESCAPE_PARAMS; break;
    case TGB_CTRL_SERIALIZED_SPINE: need = TGB_SPINE_HEADER + 5 * w->C; break;
    case TGB_CTRL_CPG_IMPEDANCE: return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_controller: use tgb_set_cpg_controller for the CPG controller");
    default: return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_controller: unknown controller kind");
    }
    if (kind == TGB_CTRL_NONE) { w->ctrlKind = TGB_CTRL_NONE; return TGB_OK; }
    if (!params || params_per_env != need) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_controller: wrong parameter count for this controller");
    if (kind == TGB_CTRL_SERIALIZED_SPINE) {
        const int rows = broadcast ? 1 : env_count;
        for (int e = 0; e < rows; ++e) {
            const double* P = params + (size_t)e * need;
            if (!(P[3] > 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_controller: update frequency must be positive");
            if (P[11] < 0.0 || P[11] > 3.0) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_controller: at most three loaded bodies");
            for (int k = 0; k < (int)P[11]; ++k)
                if (P[12 + 5 * k] < 0.0 || P[12 + 5 * k] >= (double)w->R) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_controller: loaded body out of range");
        }
    }
    const int n = w->cfg.n_env;
    if (kind == TGB_CTRL_ESCAPE_SINE && w->C < 24)
        return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_controller: the sine controller drives the 24 muscles of the SUPERball cluster map (EscapeController.cpp:245-291)");
    /* tgb_step is asynchronous on the world's stream: nothing of the controller may change under a step kernel in flight */
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    if (w->ctrlKind != kind) {
        /* one controller kind per world: switching kinds needs parameters for every environment.  Everything that can fail is done
           on temporaries; the world's controller (buffer, stride, kind) changes only once all of it succeeded. */
        if (!(env_begin == 0 && env_count == n)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_controller: a new controller kind must cover all environments");
        {
            tgb_step_args a = stepArgs(w, 1e-3, 1);
            a.ctrl_stride = need;
            a.ctrl_kind = kind;
            if (w->ops->step_smem(a, w->lanes, w->threads) > 227 * 1024) return fail(TGB_ERR_UNSUPPORTED, "controller parameters do not fit in shared memory");
        }
        double* newCtrl = nullptr;
        TGB_CUDA(cudaMalloc(&newCtrl, (size_t)n * need * sizeof(double)));
        if (kind == TGB_CTRL_ESCAPE_SINE && !w->d_sine) {
            cudaError_t e = cudaMalloc(&w->d_sine, (size_t)n * 48 * sizeof(double));
            if (e != cudaSuccess) { cudaFree(newCtrl); w->d_sine = nullptr; return cudaFail(e, "tgb_set_controller: sine state"); }
        }
        if (w->d_ctrl) cudaFree(w->d_ctrl);
        w->d_ctrl = newCtrl;
        w->ctrlStride = need;
        w->ctrlKind = kind;
    }
    if (env_count > 0) {
        /* stream-ordered behind anything still queued (the stream is idle here, see above) */
        if (broadcast) {
            std::vector<double> rows((size_t)env_count * need);
            for (int e = 0; e < env_count; ++e) std::memcpy(&rows[(size_t)e * need], params, need * sizeof(double));
            TGB_CUDA(cudaMemcpyAsync(w->d_ctrl + (size_t)env_begin * need, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice, w->stream));
            TGB_CUDA(cudaStreamSynchronize(w->stream)); /* `rows` goes out of scope */
        } else {
            TGB_CUDA(cudaMemcpyAsync(w->d_ctrl + (size_t)env_begin * need, params, (size_t)env_count * need * sizeof(double), cudaMemcpyHostToDevice, w->stream));
        }
        /* controller time restarts (EscapeController::onSetup: m_totalTime = 0); the spare word after it is cleared too: the
           sine phases of these environments start again from exact values */
        static_assert(FR_SPARE == FR_CTRLTIME + 1, "controller time and spare word are adjacent");
        TGB_CUDA(cudaMemset2DAsync(w->d_frame + (size_t)env_begin * TGB_NFR + FR_CTRLTIME, TGB_NFR * sizeof(double), 0, 2 * sizeof(double), env_count, w->stream));
        TGB_CUDA(cudaStreamSynchronize(w->stream));
    }
    w->ctrlKind = kind;
    return TGB_OK;
}

tgb_status tgb_set_cpg_controller(tgb_world* w, const tgb_cpg_desc* d, const double* params, int32_t params_per_env, int32_t env_begin,
                                  int32_t env_count, int32_t broadcast) {
    TGB_NEED_FINAL(w);
    if (!d || !params || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_cpg_controller: bad arguments");
    if (d->n_nodes != w->C || d->n_couplings < 0 || d->n_pairs < 0 || !d->coupling_begin || !d->impedance ||
        (d->n_couplings && (!d->coupling_target || !d->coupling_pair)))
        return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_cpg_controller: one CPG node per cable is required");
    if (params_per_env != 2 + 2 * d->n_pairs) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_cpg_controller: params_per_env must be 2 + 2 * n_pairs");
    if (d->score_body < 0 || d->score_body >= w->R) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_cpg_controller: score body out of range");
    if (!(d->control_time >= 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "control time is negative.");
    if (!(d->cable_control_step >= 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "Negative control step"); /* tgCPGCableControl.cpp:41-44 */
    if (d->cable_control != 0 && d->cable_control != 1) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_cpg_controller: unknown cable control");
    if (d->coupling_begin[0] != 0 || d->coupling_begin[d->n_nodes] != d->n_couplings) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_cpg_controller: bad CSR offsets");
    for (int i = 0; i < d->n_couplings; ++i)
        if (d->coupling_target[i] < 0 || d->coupling_target[i] >= d->n_nodes || d->coupling_pair[i] < 0 || d->coupling_pair[i] >= d->n_pairs)
            return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_cpg_controller: coupling index out of range");
    for (int i = 0; i < 4 * d->n_nodes; ++i)
        if (d->impedance[i] < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "impedance control gains must not be negative.");
    const int n = w->cfg.n_env;
    const bool fresh = w->ctrlKind != TGB_CTRL_CPG_IMPEDANCE;
    if (fresh && !(env_begin == 0 && env_count == n))
        return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_cpg_controller: a new controller kind must cover all environments");
    if (!fresh && w->ctrlStride != params_per_env && !(env_begin == 0 && env_count == n))
        return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_cpg_controller: a change of params_per_env must cover all environments");
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    /* static part */
    std::vector<int> topo;
    topo.insert(topo.end(), d->coupling_begin, d->coupling_begin + d->n_nodes + 1);
    topo.insert(topo.end(), d->coupling_target, d->coupling_target + d->n_couplings);
    topo.insert(topo.end(), d->coupling_pair, d->coupling_pair + d->n_couplings);
    if (w->d_cpg_topo) { cudaFree(w->d_cpg_topo); w->d_cpg_topo = nullptr; }
    if (w->d_cpg_imp) { cudaFree(w->d_cpg_imp); w->d_cpg_imp = nullptr; }
    TGB_CUDA(cudaMalloc(&w->d_cpg_topo, topo.size() * sizeof(int)));
    TGB_CUDA(cudaMemcpy(w->d_cpg_topo, topo.data(), topo.size() * sizeof(int), cudaMemcpyHostToDevice));
    TGB_CUDA(cudaMalloc(&w->d_cpg_imp, (size_t)4 * d->n_nodes * sizeof(double)));
    TGB_CUDA(cudaMemcpy(w->d_cpg_imp, d->impedance, (size_t)4 * d->n_nodes * sizeof(double), cudaMemcpyHostToDevice));
    const size_t stateDoubles = (size_t)TGB_CPG_STATE_DOUBLES(w->C);
    if (fresh || w->ctrlStride != params_per_env) {
        /* a different row length reallocates the parameter rows of every environment: the call must bring all of them */
        if (!(env_begin == 0 && env_count == n))
            return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_cpg_controller: a change of params_per_env must cover all environments");
        double* newCtrl = nullptr;
        TGB_CUDA(cudaMalloc(&newCtrl, (size_t)n * params_per_env * sizeof(double)));
        if (w->d_ctrl) cudaFree(w->d_ctrl);
        w->d_ctrl = newCtrl;
        w->ctrlStride = params_per_env;
    }
    if (!w->d_cpg_state) TGB_CUDA(cudaMalloc(&w->d_cpg_state, (size_t)n * stateDoubles * sizeof(double)));
    w->cpg = *d;
    w->cpg.coupling_begin = nullptr; w->cpg.coupling_target = nullptr; w->cpg.coupling_pair = nullptr; w->cpg.impedance = nullptr;
    w->ctrlKind = TGB_CTRL_CPG_IMPEDANCE;
    {
        /* The CPG integrator runs lanes over nodes (one node per cable) and dominates the step: unless the caller fixed the
           geometry, give an environment as many lanes as it has nodes (up to a warp) -- measured on B200: TetraSpine 16 lanes
           4.8e7 env-steps/s against 2.5e7 with 4, the 12-segment kinematic spine 32 lanes 8.4e6 against 5.4e6 with 16. */
        if (!w->lanesOverride) {
            int lanes = std::max(w->lanes, std::min(32, nextPow2(std::max(w->R, w->C))));
            while (lanes > w->lanes) {
                tgb_step_args a = stepArgs(w, 1e-3, 1);
                if (w->ops->step_smem(a, lanes, w->threads) <= 113 * 1024) break; /* keep two CTAs per SM */
                lanes >>= 1;
            }
            w->lanes = std::max(lanes, w->lanes);
        }
        tgb_step_args a = stepArgs(w, 1e-3, 1);
        if (w->ops->step_smem(a, w->lanes, w->threads) > 227 * 1024) {
            w->ctrlKind = TGB_CTRL_NONE;
            return fail(TGB_ERR_UNSUPPORTED, "CPG controller work space does not fit in shared memory");
        }
    }
    if (env_count > 0) {
        if (broadcast) {
            std::vector<double> rows((size_t)env_count * params_per_env);
            for (int e = 0; e < env_count; ++e) std::memcpy(&rows[(size_t)e * params_per_env], params, params_per_env * sizeof(double));
            TGB_CUDA(cudaMemcpy(w->d_ctrl + (size_t)env_begin * params_per_env, rows.data(), rows.size() * sizeof(double), cudaMemcpyHostToDevice));
        } else {
            TGB_CUDA(cudaMemcpy(w->d_ctrl + (size_t)env_begin * params_per_env, params, (size_t)env_count * params_per_env * sizeof(double),
                                cudaMemcpyHostToDevice));
        }
        /* a new CPGEquations: every node starts at phase 0, radius 0 (CPGNode.cpp:36-52); timers restart */
        TGB_CUDA(cudaMemset(w->d_cpg_state + (size_t)env_begin * stateDoubles, 0, (size_t)env_count * stateDoubles * sizeof(double)));
        TGB_CUDA(cudaMemset(w->d_status + env_begin, 0, (size_t)env_count * sizeof(int)));
        /* BaseSpineCPGControl::onSetup: initConditions = the scored segment's centre of mass */
        TGB_CUDA(w->ops->com(w->ioArgs(env_begin, env_count), 0, nullptr, nullptr, w->stream));
        TGB_CUDA(cudaStreamSynchronize(w->stream));
    }
    return TGB_OK;
}

tgb_status tgb_get_cpg_state(tgb_world* w, int32_t env_begin, int32_t env_count, double* out) {
    TGB_NEED_FINAL(w);
    if (!out || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_get_cpg_state: bad range or NULL buffer");
    if (w->ctrlKind != TGB_CTRL_CPG_IMPEDANCE || !w->d_cpg_state) return fail(TGB_ERR_STATE, "tgb_get_cpg_state: no CPG controller installed");
    if (env_count == 0) return TGB_OK;
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    const size_t stride = (size_t)TGB_CPG_STATE_DOUBLES(w->C);
    TGB_CUDA(cudaMemcpy2D(out, (size_t)3 * w->C * sizeof(double), w->d_cpg_state + (size_t)env_begin * stride, stride * sizeof(double),
                          (size_t)3 * w->C * sizeof(double), env_count, cudaMemcpyDeviceToHost));
    return TGB_OK;
}

tgb_status tgb_set_control_input(tgb_world* w, int32_t env_begin, int32_t env_count, const double* input, int32_t broadcast, double dt) {
    TGB_NEED_FINAL(w);
    if (!input || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_set_control_input: bad range or NULL buffer");
    if (!(dt > 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "dt is not positive.");
    if (env_count == 0 || w->C == 0) return TGB_OK;
    const size_t n = (size_t)(broadcast ? 1 : env_count) * w->C;
    for (size_t i = 0; i < n; ++i)
        if (input[i] < 0.0) return fail(TGB_ERR_INVALID_ARGUMENT, "Rest length is negative."); /* tgBasicActuator.cpp:139-142 */
    tgb_status st = w->ensureStage(n);
    if (st != TGB_OK) return st;
    TGB_CUDA(cudaMemcpyAsync(w->d_stage, input, n * sizeof(double), cudaMemcpyHostToDevice, w->stream));
    TGB_CUDA(w->ops->control_input(w->ioArgs(env_begin, env_count), w->d_stage, broadcast, dt, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_step(tgb_world* w, double dt, int32_t n_steps) {
    TGB_NEED_FINAL(w);
    if (!(dt > 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "dt is not positive"); /* tgSimulation.cpp:197-200 */
    if (n_steps < 0) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_step: negative step count");
    if (n_steps == 0) return TGB_OK;
    return launchStep(w, dt, n_steps);
}

tgb_status tgb_step_world(tgb_world* w, double dt) {
    TGB_NEED_FINAL(w);
    if (!(dt > 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "dt is not positive"); /* tgWorld.cpp:121-124 */
    return launchStep(w, dt, 1, 1);
}

tgb_status tgb_step_models(tgb_world* w, double dt) {
    TGB_NEED_FINAL(w);
    if (!(dt > 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "dt is not positive."); /* tgBasicActuator.cpp:101-104 */
    return launchStep(w, dt, 1, 2);
}

tgb_status tgb_sync(tgb_world* w) {
    TGB_NEED_FINAL(w);
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_rollout_host(tgb_world* w, const double* rod_state_in, const double* cable_rest_prev_in, double dt, int32_t n_steps,
                            double* rod_state_out, double* cable_obs_out) {
    TGB_NEED_FINAL(w);
    if (!(dt > 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "dt is not positive");
    if (n_steps < 0) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_rollout_host: negative step count");
    const int n = w->cfg.n_env;
    const size_t rodN = (size_t)n * w->R * TGB_NBS, cabInN = (size_t)n * w->C * 2, cabOutN = (size_t)n * w->C * TGB_CABLE_OBS_DOUBLES;
    const size_t stageN = rodN + cabInN + cabOutN; /* chunks in flight use disjoint staging regions */
    tgb_status st = w->ensureStage(stageN);
    if (st != TGB_OK) return st;
    /* Environments are independent, so a large batch is cut into chunks that run on their own streams: the host-to-device copy
       and packing of one chunk, the steps of another and the device-to-host copy of a third overlap (TGB_ROLLOUT_CHUNKS=1..4
       overrides the choice for batches of at least 1024 environments). */
    /* measured on B200, SUPERball, 1000 steps per call: 4096 envs 8.79e8 / 8.93e8 / 8.62e8 env-steps/s for 1 / 2 / 4 chunks */
    int chunks = n >= 16384 ? (int)tgb_world::kPipe : n >= 4096 ? 2 : 1;
    {
        const char* envChunks = std::getenv("TGB_ROLLOUT_CHUNKS");
        if (envChunks && n >= 1024) chunks = std::max(1, std::min((int)tgb_world::kPipe, std::atoi(envChunks)));
    }
    if (chunks > 1 && !w->pipeStart) {
        cudaError_t e = cudaEventCreateWithFlags(&w->pipeStart, cudaEventDisableTiming);
        for (int i = 0; i < tgb_world::kPipe && e == cudaSuccess; ++i) {
            e = cudaStreamCreateWithFlags(&w->pipe[i], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->pipeDone[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) return cudaFail(e, "tgb_rollout_host: stream/event creation");
    }
    if (chunks > 1) TGB_CUDA(cudaEventRecord(w->pipeStart, w->stream));
    const int groups = w->threads / w->lanes;
    int per = (n + chunks - 1) / chunks;
    per = (per + groups - 1) / groups * groups; /* whole CTAs per chunk */
    double* dRodStage = w->d_stage;
    double* dCabStage = w->d_stage + rodN;
    double* dObsStage = dCabStage + cabInN;
    for (int k = 0; k < chunks; ++k) {
        const int b = k * per, cnt = std::min(per, n - b);
        if (cnt <= 0) break;
        cudaStream_t s = chunks > 1 ? w->pipe[k] : w->stream;
        if (chunks > 1) TGB_CUDA(cudaStreamWaitEvent(s, w->pipeStart, 0));
        const tgb::tgb_io_args io = w->ioArgs(b, cnt);
        const size_t rodOff = (size_t)b * w->R * TGB_NBS, rodCnt = (size_t)cnt * w->R * TGB_NBS;
        if (rod_state_in) {
            TGB_CUDA(cudaMemcpyAsync(dRodStage + rodOff, rod_state_in + rodOff, rodCnt * sizeof(double), cudaMemcpyHostToDevice, s));
            TGB_CUDA(w->ops->pack_rods(io, dRodStage + rodOff, 0, s));
            TGB_CUDA(w->ops->com(io, 0, nullptr, nullptr, s));
        }
        if (cable_rest_prev_in && w->C) {
            const size_t off = (size_t)b * w->C * 2, cn = (size_t)cnt * w->C * 2;
            TGB_CUDA(cudaMemcpyAsync(dCabStage + off, cable_rest_prev_in + off, cn * sizeof(double), cudaMemcpyHostToDevice, s));
            TGB_CUDA(w->ops->pack_cables(io, dCabStage + off, 0, 0, s));
        }
        if (n_steps > 0) {
            st = launchStepRange(w, dt, n_steps, b, cnt, s, chunks > 1 ? 1 + k : 0);
            if (st != TGB_OK) return st;
        }
        if (rod_state_out) {
            TGB_CUDA(w->ops->unpack_rods(io, dRodStage + rodOff, s));
            TGB_CUDA(cudaMemcpyAsync(rod_state_out + rodOff, dRodStage + rodOff, rodCnt * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        if (cable_obs_out && w->C) {
            const size_t off = (size_t)b * w->C * TGB_CABLE_OBS_DOUBLES, cn = (size_t)cnt * w->C * TGB_CABLE_OBS_DOUBLES;
            TGB_CUDA(w->ops->cable_obs(io, dObsStage + off, s));
            TGB_CUDA(cudaMemcpyAsync(cable_obs_out + off, dObsStage + off, cn * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        if (chunks > 1) {
            TGB_CUDA(cudaEventRecord(w->pipeDone[k], s));
            TGB_CUDA(cudaStreamWaitEvent(w->stream, w->pipeDone[k], 0));
        }
    }
    if (n_steps > 0) w->time += dt * n_steps;
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_get_time(const tgb_world* w, double* out) {
    if (!w || !out) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = w->time;
    return TGB_OK;
}

tgb_status tgb_set_height_window(tgb_world* w, double lo, double hi) {
    if (!w) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL world");
    w->heightLo = lo; w->heightHi = hi;
    return TGB_OK;
}

tgb_status tgb_get_status(tgb_world* w, int32_t env_begin, int32_t env_count, int32_t* out) {
    TGB_NEED_FINAL(w);
    if (!out || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_get_status: bad range or NULL buffer");
    if (env_count == 0) return TGB_OK;
    TGB_CUDA(cudaMemcpyAsync(out, w->d_status + env_begin, (size_t)env_count * sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_mark_episode_start(tgb_world* w) {
    TGB_NEED_FINAL(w);
    TGB_CUDA(w->ops->com(w->ioArgs(0, w->cfg.n_env), 0, nullptr, nullptr, w->stream));
    TGB_CUDA(cudaMemsetAsync(w->d_status, 0, (size_t)w->cfg.n_env * sizeof(int), w->stream));
    if (w->d_max_tension) TGB_CUDA(cudaMemsetAsync(w->d_max_tension, 0, (size_t)w->cfg.n_env * std::max(1, w->C) * w->scalarBytes, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_track_max_tension(tgb_world* w, int32_t enable) {
    TGB_NEED_FINAL(w);
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    if (!enable) {
        if (w->d_max_tension) { cudaFree(w->d_max_tension); w->d_max_tension = nullptr; }
        return TGB_OK;
    }
    const size_t bytes = (size_t)w->cfg.n_env * std::max(1, w->C) * w->scalarBytes;
    if (!w->d_max_tension) TGB_CUDA(cudaMalloc(&w->d_max_tension, bytes));
    TGB_CUDA(cudaMemset(w->d_max_tension, 0, bytes));
    return TGB_OK;
}

tgb_status tgb_get_max_tension(tgb_world* w, int32_t env_begin, int32_t env_count, double* out) {
    TGB_NEED_FINAL(w);
    if (!out || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_get_max_tension: bad range or NULL buffer");
    if (!w->d_max_tension) return fail(TGB_ERR_STATE, "tgb_get_max_tension: tension history is off (tgb_track_max_tension)");
    if (env_count == 0 || w->C == 0) return TGB_OK;
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    const size_t n = (size_t)env_count * w->C;
    std::vector<unsigned char> tmp(n * w->scalarBytes);
    TGB_CUDA(cudaMemcpy(tmp.data(), (const unsigned char*)w->d_max_tension + (size_t)env_begin * w->C * w->scalarBytes, tmp.size(),
                        cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i)
        out[i] = w->scalarBytes == 4 ? (double)reinterpret_cast<const float*>(tmp.data())[i] : reinterpret_cast<const double*>(tmp.data())[i];
    return TGB_OK;
}

tgb_status tgb_get_scores(tgb_world* w, int32_t env_begin, int32_t env_count, double* scores) {
    TGB_NEED_FINAL(w);
    if (!scores || !rangeOk(w, env_begin, env_count)) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_get_scores: bad range or NULL buffer");
    if (env_count == 0) return TGB_OK;
    const size_t n = (size_t)env_count * 2;
    tgb_status st = w->ensureStage(n);
    if (st != TGB_OK) return st;
    TGB_CUDA(w->ops->com(w->ioArgs(env_begin, env_count), 1, w->d_stage, w->d_scores, w->stream));
    TGB_CUDA(cudaMemcpyAsync(scores, w->d_stage, n * sizeof(double), cudaMemcpyDeviceToHost, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    return TGB_OK;
}

tgb_status tgb_scores_device_ptr(tgb_world* w, void** ptr, size_t* bytes) {
    TGB_NEED_FINAL(w);
    if (!ptr || !bytes) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL argument");
    /* refresh the float score buffer for all environments */
    TGB_CUDA(w->ops->com(w->ioArgs(0, w->cfg.n_env), 1, nullptr, w->d_scores, w->stream));
    TGB_CUDA(cudaStreamSynchronize(w->stream));
    *ptr = w->d_scores;
    *bytes = (size_t)w->cfg.n_env * 2 * sizeof(float);
    return TGB_OK;
}

tgb_status tgb_scores_refresh_async(tgb_world* w, void** ptr, size_t* bytes) {
    TGB_NEED_FINAL(w);
    if (!ptr || !bytes) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL argument");
    TGB_CUDA(w->ops->com(w->ioArgs(0, w->cfg.n_env), 1, nullptr, w->d_scores, w->stream));
    *ptr = w->d_scores;
    *bytes = (size_t)w->cfg.n_env * 2 * sizeof(float);
    return TGB_OK;
}

tgb_status tgb_step_timed(tgb_world* w, double dt, int32_t n_steps, int32_t reps, float* ms_out) {
    TGB_NEED_FINAL(w);
    if (!(dt > 0.0)) return fail(TGB_ERR_INVALID_ARGUMENT, "dt is not positive");
    if (n_steps < 1 || reps < 1 || !ms_out) return fail(TGB_ERR_INVALID_ARGUMENT, "tgb_step_timed: bad arguments");
    for (int r = 0; r < reps; ++r) {
        TGB_CUDA(cudaEventRecord(w->ev0, w->stream));
        tgb_status st = launchStep(w, dt, n_steps);
        if (st != TGB_OK) return st;
        TGB_CUDA(cudaEventRecord(w->ev1, w->stream));
        TGB_CUDA(cudaEventSynchronize(w->ev1));
        TGB_CUDA(cudaEventElapsedTime(&ms_out[r], w->ev0, w->ev1));
    }
    return TGB_OK;
}

tgb_status tgb_launch_count(const tgb_world* w, int64_t* out) {
    if (!w || !out) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = w->launches;
    return TGB_OK;
}

tgb_status tgb_launch_geometry(const tgb_world* w, int32_t* lanes_per_env, int32_t* block, int32_t* grid) {
    if (!w) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL world");
    if (!w->finalized) return fail(TGB_ERR_STATE, "world is not finalized");
    const bool wide = wideLayout(w);
    const int groups = wide ? 32 / w->lanes : w->threads / w->lanes;
    if (lanes_per_env) *lanes_per_env = w->lanes;
    if (block) *block = wide ? 128 : w->threads;
    if (grid) *grid = (w->cfg.n_env + groups - 1) / groups;
    return TGB_OK;
}

tgb_status tgb_state_bytes_per_env(const tgb_world* w, size_t* out) {
    if (!w || !out) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!w->finalized) return fail(TGB_ERR_STATE, "world is not finalized");
    *out = (size_t)w->stride * w->scalarBytes + TGB_NFR * sizeof(double);
    return TGB_OK;
}

tgb_status tgb_stream_handle(tgb_world* w, void** stream) {
    if (!w || !stream) return fail(TGB_ERR_INVALID_ARGUMENT, "NULL argument");
    *stream = (void*)w->stream;
    return TGB_OK;
}

} /* extern "C" */
