/*
 * tgb_ccable.cuh -- contact cables on the device: tgBulletContactSpringCable::step (src/core/tgBulletContactSpringCable.cpp:97-125)
 * with its sliding anchors (src/core/tgBulletSpringCableAnchor.cpp:89-334) for every environment, run after the step kernel's models
 * half (tgBasicActuator::step of a contact cable is the cable's step, src/core/tgBasicActuator.cpp:227-241).
 *
 * The terms of the restatement are in include/ntrtsim_b200.h ("Contact cables"): Bullet's ghost object and persistent manifolds
 * become, per step, the closest-point pairs of the cable's segments with the capsules / spheres of every body.
 *
 * Mapping: one thread per environment.  The anchor list of a cable is a short, ordered, data-dependent list (insertions, sideways
 * slides, removals until four passes change nothing): there is no lane-parallel form of it worth the name, but environments are
 * independent, so 4096 of them fill the lanes.  A thread walks its environment's contact cables one after the other and adds the
 * anchor impulses to the bodies' velocities in the environment record in a fixed order: results do not depend on the launch geometry.
 * All arithmetic is double (the cable is thin -- 1e-3 against rods of 0.25 -- and the anchor tests compare lengths with sums of
 * lengths); the record is read and written in the world's scalar type.
 *
 * HBM: cc_state double[n_env][n_cc][TGB_CC_STATE] = rest length, previous length, velocity, number of sliding anchors, preferred length, then
 * TGB_CCABLE_MAX_SLIDING x (body, local position 3, contact normal 3); cc_model double[n_cc][TGB_CC_MODEL].
 */
#pragma once
#include "../../include/ntrtsim_b200.h"
#include "tgb_dev.h"
#include "tgb_math.cuh"

namespace tgb {

enum { CCM_A = 0, CCM_B, CCM_LA = 2, CCM_LB = 5, CCM_K = 8, CCM_C, CCM_THICK, CCM_RES, CCM_MAXTENS, CCM_TVEL, CCM_MINAL, CCM_MINRL, TGB_CC_MODEL = 16 };
enum { CCS_REST = 0, CCS_PREV, CCS_VEL, CCS_N, CCS_PREF, CCS_ANCHOR = 5, TGB_CC_ANCHOR = 7, TGB_CC_STATE = CCS_ANCHOR + TGB_CCABLE_MAX_SLIDING * TGB_CC_ANCHOR };
enum { TGB_CC_MAX_POINTS = 64, TGB_CC_CHAIN = TGB_CCABLE_MAX_SLIDING + 2 };

struct tgb_ccable_args {
    void* state;
    const double* frame;
    const void* model_t;
    const int* model_i;
    double* cc_state;
    const double* cc_model;
    int n_env, env_begin, env_count, R, S, stride, n_cc;
    double dt;
};

typedef V3<double> D3;

struct CcAnchor { int body; D3 local, normal; }; /* normal: in the body frame (USE_BASIS), turned into the world frame on use */
struct CcPoint { int body; D3 pos, normal; double dist; };

/* the environment's record seen in double.  Bodies do not move during a cable's step (only velocities change), so their poses are
   converted once per launch into a per-thread cache (local memory, L1-resident: 96 bytes per body) instead of once per use -- the
   anchor passes ask for world positions some hundred times per step. */
enum { TGB_CC_MAX_BODIES = 32 }; /* tgb_finalize refuses more than 32 bodies per environment */
template <typename T>
struct CcEnv {
    T* rec;
    const T* mt;
    const int* mi;
    int R, S;
    D3 cpos[TGB_CC_MAX_BODIES];
    M3<double> cbasis[TGB_CC_MAX_BODIES];
    __device__ void cachePoses() {
        for (int b = 0; b < R; ++b) {
            const T* s = rec + b * TGB_NBS;
            cpos[b] = D3((double)s[0], (double)s[1], (double)s[2]);
            cbasis[b] = M3<double>::fromQuat(Q4<double>((double)s[3], (double)s[4], (double)s[5], (double)s[6]));
        }
    }
    __device__ const D3& pos(int b) const { return cpos[b]; }
    __device__ const M3<double>& basis(int b) const { return cbasis[b]; }
    __device__ D3 world(int b, const D3& local) const { return cbasis[b] * local + cpos[b]; }
    __device__ D3 local(int b, const D3& w) const { return cbasis[b].tmul(w - cpos[b]); }
    /* tgBulletSpringCableAnchor::getContactNormal (:202-214, USE_BASIS): the anchor's body-frame normal in the world frame, normalised */
    __device__ D3 worldNormal(int b, const D3& nl) const {
        const D3 n = cbasis[b] * nl;
        const double l = n.length();
        return l > 0.0 ? n / l : D3(0.0, 0.0, 0.0);
    }
    __device__ double bodyField(int f, int b) const { return (double)mt[f * R + b]; }
    __device__ double shapeField(int f, int s) const { return (double)mt[TGB_NBF * R + f * S + s]; }
    __device__ int bodyInt(int f, int b) const { return mi[f * R + b]; }
    /* btRigidBody::applyImpulse */
    __device__ void applyImpulse(int b, const D3& J, const D3& rel) {
        const double im = bodyField(BF_INVMASS, b);
        if (im == 0.0) return;
        T* s = rec + b * TGB_NBS;
        const M3<double> Iinv = cbasis[b].conjugateDiag(D3(bodyField(BF_INVIX, b), bodyField(BF_INVIY, b), bodyField(BF_INVIZ, b)));
        const D3 dw = Iinv * rel.cross(J);
        s[7] = (T)((double)s[7] + J.x * im); s[8] = (T)((double)s[8] + J.y * im); s[9] = (T)((double)s[9] + J.z * im);
        s[10] = (T)((double)s[10] + dw.x); s[11] = (T)((double)s[11] + dw.y); s[12] = (T)((double)s[12] + dw.z);
    }
};

struct CcCable {
    int a, b, n;
    D3 la, lb;
    CcAnchor s[TGB_CCABLE_MAX_SLIDING];
};

template <typename T>
__device__ int ccChain(const CcEnv<T>& E, const CcCable& c, D3* P) {
    P[0] = E.world(c.a, c.la);
    for (int i = 0; i < c.n; ++i) P[i + 1] = E.world(c.s[i].body, c.s[i].local);
    P[c.n + 1] = E.world(c.b, c.lb);
    return c.n + 2;
}
__device__ inline double ccLength(const D3* P, int n) {
    double L = 0.0;
    for (int i = 0; i + 1 < n; ++i) L += (P[i] - P[i + 1]).length();
    return L;
}
__device__ inline void ccErase(CcCable& c, int i) {
    for (int k = i; k + 1 < c.n; ++k) c.s[k] = c.s[k + 1];
    --c.n;
}
/* the manifold point of `body` nearest to p (compared by squared distance; one square root at the end) */
__device__ inline double ccNearest(const CcPoint* pts, int np, int body, const D3& p, D3& normal, D3& pos) {
    double best2 = INFINITY;
    int bi = -1;
    for (int i = 0; i < np; ++i) {
        if (pts[i].body != body) continue;
        const double d2 = (pts[i].pos - p).length2();
        if (d2 < best2) { best2 = d2; bi = i; }
    }
    if (bi < 0) return INFINITY;
    normal = pts[bi].normal; pos = pts[bi].pos;
    return sqrt(best2);
}
/* tgBulletSpringCableAnchor::getManifoldDistance, :279-334 */
template <typename T>
__device__ double ccManifoldDistance(const CcEnv<T>& E, const CcPoint* pts, int np, const CcAnchor& q, int body, D3& normalOut) {
    normalOut = E.worldNormal(q.body, q.normal);
    if (q.body != body) return INFINITY;
    D3 n, pos;
    const double len = ccNearest(pts, np, body, E.world(q.body, q.local), n, pos);
    if (len < 0.1) normalOut = n;
    return len;
}
/* tgBulletSpringCableAnchor::updateManifold, :232-277 */
template <typename T>
__device__ bool ccUpdateManifold(const CcEnv<T>& E, const CcPoint* pts, int np, const CcAnchor& q, int body) {
    if (q.body != body) return false;
    D3 nn;
    ccManifoldDistance(E, pts, np, q, body, nn);
    return !((nn + E.worldNormal(q.body, q.normal)).length() < 0.5);
}
/* findNearestPastAnchor, :799-937, restated: the segment of the chain nearest to pos */
__device__ inline int ccNearestSegment(const D3* P, int n, const D3& pos) {
    int best = 0;
    double bestD2 = INFINITY;
    for (int j = 0; j + 1 < n; ++j) {
        const D3 e = P[j + 1] - P[j];
        const double e2 = e.length2();
        double t = e2 > 0.0 ? (pos - P[j]).dot(e) / e2 : 0.0;
        t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
        const double d2 = (P[j] + e * t - pos).length2();
        if (d2 < bestD2) { bestD2 = d2; best = j; }
    }
    return best;
}
/* tgBulletSpringCableAnchor::setWorldPosition, :89-184 */
template <typename T>
__device__ bool ccSetWorldPosition(const CcEnv<T>& E, const CcPoint* pts, int np, CcAnchor& q, const D3& newPos) {
    const D3 current = E.world(q.body, q.local);
    const D3 own = E.worldNormal(q.body, q.normal);
    D3 newNormal = own, contactPos = current;
    const double length = ccNearest(pts, np, q.body, newPos, newNormal, contactPos);
    if (length < 0.1) {
        q.local = E.local(q.body, newPos);
        return !((newNormal + own).length() < 0.5);
    }
    return (current - contactPos).length() <= 0.1;
}
/* updateAnchorPositions, :484-561.  P = the chain, kept current: only the moved anchor's entry changes (recomputed from its new local
   position exactly as ccChain would), a removed anchor's entry is closed up. */
template <typename T>
__device__ int ccUpdateAnchorPositions(const CcEnv<T>& E, const CcPoint* pts, int np, CcCable& c, D3* P) {
    int numPruned = 0, i = 0;
    while (i < c.n) {
        const D3 back = P[i], current = P[i + 1], forward = P[i + 2];
        const D3 lineA = forward - current, lineB = back - current;
        const D3 n = E.worldNormal(c.s[i].body, c.s[i].normal);
        D3 tangentMove(0.0, 0.0, 0.0);
        const D3 along = lineB - lineA;
        const double al = along.length();
        if (al > 0.0) {
            const D3 td = (along / al).cross(n);
            const double tl = td.length();
            if (tl > 0.0) {
                const D3 tangentDir = td / tl;
                tangentMove = tangentDir * (lineB + lineA).dot(tangentDir);
            }
        }
        const D3 newPos = current + tangentMove;
        if (!ccSetWorldPosition(E, pts, np, c.s[i], newPos)) {
            ccErase(c, i);
            for (int k = i + 1; k < c.n + 2; ++k) P[k] = P[k + 1];
            ++numPruned;
        } else {
            P[i + 1] = E.world(c.s[i].body, c.s[i].local);
            ++i;
        }
    }
    return numPruned;
}

/* closest points of two segments given by centre, unit direction and half length (the routine of the body pairs, tgb_pairs.cuh) */
__device__ inline void ccClosest(const D3& pa, const D3& da, double ha, const D3& pb, const D3& db, double hb, double& sOut, double& tOut) {
    const D3 r = pa - pb;
    const double b = da.dot(db), c = da.dot(r), f = db.dot(r);
    const double denom = 1.0 - b * b;
    double s = 0.0, t = 0.0;
    if (ha > 0.0 && denom > 1e-12) { s = (b * f - c) / denom; s = s < -ha ? -ha : (s > ha ? ha : s); }
    t = b * s + f;
    if (t < -hb) { t = -hb; s = b * t - c; s = s < -ha ? -ha : (s > ha ? ha : s); }
    else if (t > hb) { t = hb; s = b * t - c; s = s < -ha ? -ha : (s > ha ? ha : s); }
    sOut = s; tOut = t;
}

template <typename T>
__global__ void __launch_bounds__(32) ccable_step_kernel(tgb_ccable_args a) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.env_count) return;
    const int env = a.env_begin + e;
    CcEnv<T> E;
    E.rec = reinterpret_cast<T*>(a.state) + (size_t)env * a.stride;
    E.mt = reinterpret_cast<const T*>(a.model_t);
    E.mi = a.model_i;
    E.R = a.R; E.S = a.S;
    E.cachePoses();
    const double dt = a.dt;
    const double kBreaking = 0.02; /* gContactBreakingThreshold */
    CcPoint pts[TGB_CC_MAX_POINTS];
    D3 P[TGB_CC_CHAIN];
    for (int ci = 0; ci < a.n_cc; ++ci) {
        const double* cm = a.cc_model + (size_t)ci * TGB_CC_MODEL;
        double* cs = a.cc_state + ((size_t)env * a.n_cc + ci) * TGB_CC_STATE;
        CcCable c;
        c.a = (int)cm[CCM_A]; c.b = (int)cm[CCM_B];
        c.la = D3(cm[CCM_LA], cm[CCM_LA + 1], cm[CCM_LA + 2]);
        c.lb = D3(cm[CCM_LB], cm[CCM_LB + 1], cm[CCM_LB + 2]);
        const double k = cm[CCM_K], damp = cm[CCM_C], thickness = cm[CCM_THICK], resolution = cm[CCM_RES];
        const double rest = cs[CCS_REST], prev = cs[CCS_PREV];
        c.n = (int)cs[CCS_N];
        for (int i = 0; i < c.n; ++i) {
            const double* q = cs + CCS_ANCHOR + i * TGB_CC_ANCHOR;
            c.s[i].body = (int)q[0];
            c.s[i].local = D3(q[1], q[2], q[3]);
            c.s[i].normal = D3(q[4], q[5], q[6]);
        }
        /* updateManifolds: this step's contact points */
        int nP = ccChain(E, c, P), np = 0;
        for (int i = 0; i + 1 < nP; ++i) {
            const D3 ed = P[i + 1] - P[i];
            const double len = ed.length();
            const D3 gp = (P[i] + P[i + 1]) * 0.5;
            const D3 gd = len > 1e-12 ? ed / len : D3(0.0, 1.0, 0.0);
            const double gh = len > 1e-12 ? 0.5 * len : 0.0;
            for (int bi = 0; bi < a.R; ++bi) {
                const int s0 = E.bodyInt(BI_SHAPE_BEGIN, bi), sn = E.bodyInt(BI_SHAPE_COUNT, bi);
                if (sn <= 0) continue;
                const M3<double>& B = E.basis(bi);
                const D3 x = E.pos(bi);
                for (int si = s0; si < s0 + sn; ++si) {
                    if (E.shapeField(SF_KIND, si) == (double)TGB_SHAPE_BOX) continue;
                    const D3 sp(E.shapeField(SF_POSX, si), E.shapeField(SF_POSY, si), E.shapeField(SF_POSZ, si));
                    const D3 r1(E.shapeField(SF_ROT0 + 1, si), E.shapeField(SF_ROT0 + 4, si), E.shapeField(SF_ROT0 + 7, si));
                    const D3 bp = B * sp + x, bd = B * r1;
                    const double br = E.shapeField(SF_RADIUS, si), bh = E.shapeField(SF_SEGH, si);
                    double s, t;
                    ccClosest(gp, gd, gh, bp, bd, bh, s, t);
                    const D3 ca = gp + gd * s, cb = bp + bd * t;
                    const D3 dv = ca - cb;
                    const double d = dv.length();
                    if (!(d > 1e-9)) continue;
                    const double dist = d - br - thickness;
                    if (dist > kBreaking) continue;
                    if (np >= TGB_CC_MAX_POINTS) continue;
                    const D3 u = dv / d;
                    pts[np].body = bi; pts[np].pos = cb + u * br; pts[np].normal = -u; pts[np].dist = dist;
                    ++np;
                }
            }
        }
        /* updateManifolds + updateAnchorList (:196-482): proposals of the penetrating points, one after the other.  A proposal changes
           the cable only if it passes every test (the tests themselves change nothing here: an anchor's manifold is its body's point
           set), so they run cheapest first -- the cap, the two sides of the contact plane, the neighbours' normals -- and the merge
           test with its manifold distances last; the chain is recomputed only after an insertion. */
        nP = ccChain(E, c, P);
        for (int pi = 0; pi < np; ++pi) {
            if (c.n >= TGB_CCABLE_MAX_SLIDING) break;
            const CcPoint& m = pts[pi];
            if (!(m.dist < 0.0)) continue;
            const int j = ccNearestSegment(P, nP, m.pos);
            const D3 lineA = P[j + 1] - m.pos, lineB = P[j] - m.pos;
            if (lineA.dot(m.normal) < 0.0 || lineB.dot(m.normal) < 0.0) continue;
            const bool backSliding = j >= 1, fwdSliding = j + 1 <= c.n;
            const int backBody = backSliding ? c.s[j - 1].body : c.a, fwdBody = fwdSliding ? c.s[j].body : c.b;
            if ((backSliding && m.body == backBody && E.worldNormal(c.s[j - 1].body, c.s[j - 1].normal).dot(m.normal) < 0.0) ||
                (fwdSliding && m.body == fwdBody && E.worldNormal(c.s[j].body, c.s[j].normal).dot(m.normal) < 0.0)) continue;
            const double lengthA = lineA.length(), lengthB = lineB.length();
            const bool nearBack = lengthB <= resolution && m.body == backBody && backSliding;
            const bool nearFwd = lengthA <= resolution && m.body == fwdBody && fwdSliding;
            if (nearBack || nearFwd) {
                D3 tmp;
                const double mDistB = backSliding ? ccManifoldDistance(E, pts, np, c.s[j - 1], m.body, tmp) : INFINITY;
                const double mDistA = fwdSliding ? ccManifoldDistance(E, pts, np, c.s[j], m.body, tmp) : INFINITY;
                bool del = false;
                if (nearBack && mDistB < mDistA && ccUpdateManifold(E, pts, np, c.s[j - 1], m.body)) del = true;
                if (nearFwd && (!del || mDistA < mDistB) && ccUpdateManifold(E, pts, np, c.s[j], m.body)) del = true;
                if (del) continue;
            }
            for (int q = c.n; q > j; --q) c.s[q] = c.s[q - 1];
            c.s[j].body = m.body; c.s[j].local = E.local(m.body, m.pos); c.s[j].normal = E.basis(m.body).tmul(m.normal); /* basis^-1 * cn */
            ++c.n;
            nP = ccChain(E, c, P);
            if (ccLength(P, nP) > prev + 2.0 * resolution) { ccErase(c, j); nP = ccChain(E, c, P); }
        }
        /* pruneAnchors, :563-673 */
        int numPruned = 0, passes = 0;
        for (int i = 0; i < c.n;) {
            D3 n, pos;
            if (ccNearest(pts, np, c.s[i].body, E.world(c.s[i].body, c.s[i].local), n, pos) == INFINITY) { ccErase(c, i); ++numPruned; }
            else ++i;
        }
        nP = ccChain(E, c, P);
        while (numPruned > 0 || passes <= 3) {
            numPruned = ccUpdateAnchorPositions(E, pts, np, c, P);
            if (numPruned == 0) {
                int i = 0;
                while (i < c.n) {
                    const D3 lineA = P[i + 2] - P[i + 1], lineB = P[i] - P[i + 1];
                    const D3 n = E.worldNormal(c.s[i].body, c.s[i].normal);
                    double nv1, nv2;
                    const double half = resolution / 2.0;
                    if (lineA.length2() < half * half || lineB.length2() < half * half) nv1 = nv2 = -1.0;
                    else { nv1 = lineA.dot(n); nv2 = lineB.dot(n); }
                    if (nv1 < 0.0 || nv2 < 0.0) {
                        ccErase(c, i);
                        for (int k = i + 1; k < c.n + 2; ++k) P[k] = P[k + 1];
                        ++numPruned;
                    } else ++i;
                }
            }
            if (numPruned == 0) ++passes; else passes = 0;
            if (c.n == 0) break; /* nothing left to slide or remove: the remaining passes are empty */
        }
        /* calculateAndApplyForce, :127-194 */
        nP = ccChain(E, c, P);
        const double currLength = ccLength(P, nP);
        double tension = (currLength - rest) * k;
        if (tension < 0.0) tension = 0.0;
        const double velocity = (currLength - prev) / dt;
        double damping = damp * velocity;
        if (fabs(tension) * 1.0 < fabs(damping)) damping = (damping > 0.0 ? tension * 1.0 : -tension * 1.0);
        const double magnitude = tension + damping;
        for (int i = 0; i < nP; ++i) {
            D3 force(0.0, 0.0, 0.0);
            if (i == 0) { const D3 d = P[1] - P[0]; const double l = d.length(); if (l > 0.0) force = (d / l) * magnitude; }
            else if (i == nP - 1) { const D3 d = P[i] - P[i - 1]; const double l = d.length(); if (l > 0.0) force = -(d / l) * magnitude; }
            else {
                const D3 first = P[i + 1] - P[i], second = P[i - 1] - P[i];
                const double l1 = first.length(), l2 = second.length();
                D3 dir(0.0, 0.0, 0.0);
                if (l1 > 0.0) dir += first / l1;
                if (l2 > 0.0) dir += second / l2;
                force = dir * magnitude;
            }
            const int body = i == 0 ? c.a : (i == nP - 1 ? c.b : c.s[i - 1].body);
            E.applyImpulse(body, force * dt, P[i] - E.pos(body));
        }
        cs[CCS_PREV] = currLength;
        cs[CCS_VEL] = velocity;
        cs[CCS_N] = (double)c.n;
        for (int i = 0; i < c.n; ++i) {
            double* q = cs + CCS_ANCHOR + i * TGB_CC_ANCHOR;
            q[0] = (double)c.s[i].body;
            q[1] = c.s[i].local.x; q[2] = c.s[i].local.y; q[3] = c.s[i].local.z;
            q[4] = c.s[i].normal.x; q[5] = c.s[i].normal.y; q[6] = c.s[i].normal.z;
        }
    }
}

/* tgBasicActuator::setControlInput(input, dt) -> moveMotors (src/core/tgBasicActuator.cpp:149-225) for every contact cable; NaN input =
   leave the cable alone.  One thread per (environment, cable). */
template <typename T>
__global__ void ccable_control_kernel(tgb_ccable_args a, const double* __restrict__ input, int broadcast) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.env_count * a.n_cc) return;
    const int e = i / a.n_cc, ci = i - e * a.n_cc;
    const int env = a.env_begin + e;
    const double in = input[(size_t)(broadcast ? 0 : e) * a.n_cc + ci];
    if (!(in >= 0.0)) return;
    CcEnv<T> E;
    E.rec = reinterpret_cast<T*>(a.state) + (size_t)env * a.stride;
    E.mt = reinterpret_cast<const T*>(a.model_t);
    E.mi = a.model_i;
    E.R = a.R; E.S = a.S;
    E.cachePoses();
    const double* cm = a.cc_model + (size_t)ci * TGB_CC_MODEL;
    double* cs = a.cc_state + ((size_t)env * a.n_cc + ci) * TGB_CC_STATE;
    CcCable c;
    c.a = (int)cm[CCM_A]; c.b = (int)cm[CCM_B];
    c.la = D3(cm[CCM_LA], cm[CCM_LA + 1], cm[CCM_LA + 2]);
    c.lb = D3(cm[CCM_LB], cm[CCM_LB + 1], cm[CCM_LB + 2]);
    c.n = (int)cs[CCS_N];
    for (int q = 0; q < c.n; ++q) {
        const double* s = cs + CCS_ANCHOR + q * TGB_CC_ANCHOR;
        c.s[q].body = (int)s[0];
        c.s[q].local = D3(s[1], s[2], s[3]);
    }
    D3 P[TGB_CC_CHAIN];
    const int nP = ccChain(E, c, P);
    const double actual = ccLength(P, nP);
    const double k = cm[CCM_K], maxTens = cm[CCM_MAXTENS];
    double pref = in, rest = cs[CCS_REST];
    const double stepSize = cm[CCM_TVEL] * a.dt;
    if ((actual - pref) * k > maxTens) pref = actual - maxTens / k;
    const double diff = pref - rest;
    const double fabsDiff = fabs(diff);
    if ((actual > cm[CCM_MINAL]) || (diff > 0.0)) {
        if (fabsDiff > stepSize) rest += (diff / fabsDiff) * stepSize;
        else rest += diff;
    }
    rest = (rest > cm[CCM_MINRL]) ? rest : cm[CCM_MINRL];
    cs[CCS_REST] = rest;
    cs[CCS_PREF] = pref;
}

/* observations [env_count][n_cc][5] and, with cable >= 0, that cable's sliding anchors [env_count][MAX][7] */
template <typename T>
__global__ void ccable_obs_kernel(tgb_ccable_args a, double* __restrict__ obs, int cable, double* __restrict__ anchors) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.env_count * a.n_cc) return;
    const int e = i / a.n_cc, ci = i - e * a.n_cc;
    const int env = a.env_begin + e;
    CcEnv<T> E;
    E.rec = reinterpret_cast<T*>(a.state) + (size_t)env * a.stride;
    E.mt = reinterpret_cast<const T*>(a.model_t);
    E.mi = a.model_i;
    E.R = a.R; E.S = a.S;
    E.cachePoses();
    const double* cm = a.cc_model + (size_t)ci * TGB_CC_MODEL;
    const double* cs = a.cc_state + ((size_t)env * a.n_cc + ci) * TGB_CC_STATE;
    CcCable c;
    c.a = (int)cm[CCM_A]; c.b = (int)cm[CCM_B];
    c.la = D3(cm[CCM_LA], cm[CCM_LA + 1], cm[CCM_LA + 2]);
    c.lb = D3(cm[CCM_LB], cm[CCM_LB + 1], cm[CCM_LB + 2]);
    c.n = (int)cs[CCS_N];
    for (int q = 0; q < c.n; ++q) {
        const double* s = cs + CCS_ANCHOR + q * TGB_CC_ANCHOR;
        c.s[q].body = (int)s[0];
        c.s[q].local = D3(s[1], s[2], s[3]);
        c.s[q].normal = D3(s[4], s[5], s[6]);
    }
    D3 P[TGB_CC_CHAIN];
    const int nP = ccChain(E, c, P);
    const double L = ccLength(P, nP);
    if (obs) {
        double* o = obs + (size_t)i * TGB_CCABLE_OBS_DOUBLES;
        double tension = (L - cs[CCS_REST]) * cm[CCM_K];
        o[0] = cs[CCS_REST]; o[1] = L; o[2] = tension < 0.0 ? 0.0 : tension; o[3] = cs[CCS_VEL]; o[4] = (double)c.n;
    }
    if (anchors && ci == cable) {
        const double* fr = a.frame + (size_t)env * TGB_NFR;
        double* o = anchors + (size_t)e * TGB_CCABLE_MAX_SLIDING * TGB_CCABLE_ANCHOR_DOUBLES;
        for (int q = 0; q < TGB_CCABLE_MAX_SLIDING; ++q) {
            double* r = o + q * TGB_CCABLE_ANCHOR_DOUBLES;
            for (int k = 0; k < TGB_CCABLE_ANCHOR_DOUBLES; ++k) r[k] = 0.0;
            r[0] = -1.0;
            if (q >= c.n) continue;
            r[0] = (double)c.s[q].body;
            r[1] = fr[FR_O] + P[q + 1].x; r[2] = fr[FR_O + 1] + P[q + 1].y; r[3] = fr[FR_O + 2] + P[q + 1].z;
            const D3 wn = E.worldNormal(c.s[q].body, c.s[q].normal);
            r[4] = wn.x; r[5] = wn.y; r[6] = wn.z;
        }
    }
}

} // namespace tgb
