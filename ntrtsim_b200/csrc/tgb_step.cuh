/*
 * tgb_step.cuh -- the fused batched step kernel: n_steps x tgSimulation::step(dt) for every environment.
 *
 * Reference path replaced (paths relative to the NTRTsim tree):
 *   tgSimulation::step                 src/core/tgSimulation.cpp:193-225     world step first, then the models
 *   tgWorldBulletPhysicsImpl::step     src/core/tgWorldBulletPhysicsImpl.cpp:209-221 -> Bullet stepSimulation(dt,1,dt)
 *   tgBasicActuator::step / moveMotors src/core/tgBasicActuator.cpp:96-113,170-225
 *   tgKinematicActuator                src/core/tgKinematicActuator.cpp:135-225
 *   tgBulletSpringCable                src/core/tgBulletSpringCable.cpp:90-141
 *   anchors                            src/core/tgBulletSpringCableAnchor.cpp:76-87
 *   controllers                        src/controllers/tgTensionController.cpp:50-88,
 *                                      src/examples/craterEscape/EscapeController.cpp:100-117,245-291,
 *                                      src/examples/motorModel/tsTestRig.cpp:170-199
 *
 * Mapping.  A group of LPE consecutive lanes of one warp owns one environment for the whole launch.  Lane j of the
 * group is body j (rigid-body state lives in its registers across all n_steps) and cables j, j+LPE, ...  Per step:
 *   cable phase   lanes over cables : controller -> motors -> Hooke + damping force from the published body poses;
 *                                     the impulse J = F dt of each cable goes to shared memory
 *   body phase    lanes over bodies : gather the impulses of the incident cables in cable order (static incidence
 *                                     list; deterministic, no atomics), then the whole world step of that body:
 *                                     gravity + gyroscopic term, ground contact manifold, projected Gauss-Seidel,
 *                                     transform integration; publish the new pose
 * Bodies interact through cables and, when two of them come within the contact band, through rows between two bodies
 * (tgb_pairs.cuh).  Without such a row -- the common case, checked every step -- the ground is the only partner, the contact
 * solve of each body is independent and never leaves its lane.  Two __syncwarp() per step, no block-level barrier inside
 * the step loop.
 * HBM is touched only at launch boundaries: the CTA's environment records are one contiguous span that is moved
 * with 128-bit accesses into shared memory and back.
 *
 * No tensor cores: nothing here is a dense contraction.
 */
#pragma once
#include "../../include/ntrtsim_b200.h"
#include "tgb_dev.h"
#include "tgb_math.cuh"
#include "tgb_pairs.cuh"

#include <type_traits>

#ifndef TGB_MATRIX_PGS
#define TGB_MATRIX_PGS 1
#endif
#ifndef TGB_SINGLE_CONTACT_PGS
#define TGB_SINGLE_CONTACT_PGS 1
#endif
#ifndef TGB_EXCHANGE_UNION
#define TGB_EXCHANGE_UNION 1
#endif
#ifndef TGB_SINE_REC
#define TGB_SINE_REC 4
#endif
#ifndef TGB_VEC_CS
#define TGB_VEC_CS 1
#endif
#ifndef TGB_VEC_CREC
#define TGB_VEC_CREC 1
#endif
#ifndef TGB_FACTORED_IINV
#define TGB_FACTORED_IINV 1 /* single precision: apply the world inverse inertia as B (invI o (B^T v)) instead of holding the matrix */
#endif
#ifndef TGB_COOP_PAIRS
#define TGB_COOP_PAIRS 1 /* bodies of several shapes: pairs of shapes spread over the group's lanes (pairNearCompoundGroup) */
#endif
#ifndef TGB_MATRIX4
#define TGB_MATRIX4 0 /* three or four contact points of a body: matrix-form PGS in registers (solveContactsMatrix4).  Measured on B200,
                         same box: SUPERball 6.05e8 without / 5.89e8 with, TetraSpine 2.35e7 / 1.89e7 -- forming the 36 couplings and the
                         out-of-line call cost more than the shorter row visits save: off */
#endif
#ifndef TGB_CABLE_UNROLL
#define TGB_CABLE_UNROLL 0 /* 0 = automatic */
#endif

namespace tgb {

/* Dormand-Prince 5(4): row s = the coefficients of k1 .. k(s+1) in the state handed to stage s + 2; the last row is the fifth-order
   solution (no k2 term) */
static __constant__ double tgb_dopri_a[6][6] = {
    {1.0 / 5.0, 0.0, 0.0, 0.0, 0.0, 0.0},
    {3.0 / 40.0, 9.0 / 40.0, 0.0, 0.0, 0.0, 0.0},
    {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0.0, 0.0, 0.0},
    {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0.0, 0.0},
    {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0.0},
    {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}};

template <typename T> struct Vec16;
template <> struct Vec16<float> { typedef float4 type; enum { N = 4 }; };
template <> struct Vec16<double> { typedef double2 type; enum { N = 2 }; };

/* Bullet 2.82 btContactSolverInfo defaults reached through tgWorldBulletPhysicsImpl.cpp:134-146 (left untouched). */
template <typename T> struct SolverConst {
    static TGB_D T erp() { return T(0.2); }
    static TGB_D T erp2() { return T(0.8); }
    static TGB_D T splitThreshold() { return T(-0.04); }
    static TGB_D T turnErp() { return T(0.1); }
    static TGB_D T maxGyro() { return T(100.0); }
    static TGB_D T breakingThreshold() { return T(0.02); }
    static TGB_D T angularMotionThreshold() { return T(0.5 * 1.57079632679489661923); }
};

/* Rigid-body state held in the registers of the lane that owns the body. */
template <typename T>
struct BodyReg {
    V3<T> x;      /* position relative to the env frame origin */
    Q4<T> q;
    V3<T> v;      /* linear velocity relative to the env frame velocity */
    V3<T> w;
    M3<T> B;      /* basis = fromQuat(q) */
    M3<T> Iinv;   /* world inverse inertia */
};

/* btTransformUtil::integrateTransform on (x, q) */
template <typename T>
TGB_D void integrateTransform(V3<T>& x, Q4<T>& q, const V3<T>& linvel, const V3<T>& angvel, T dt) {
    x = x + linvel * dt;
    V3<T> axis;
    T fAngle = angvel.length();
    if (fAngle * dt > SolverConst<T>::angularMotionThreshold()) fAngle = SolverConst<T>::angularMotionThreshold() / dt;
    if (fAngle < T(0.001)) {
        axis = angvel * (T(0.5) * dt - (dt * dt * dt) * T(0.020833333333) * fAngle * fAngle);
    } else {
        axis = angvel * (tsin(T(0.5) * fAngle * dt) / fAngle);
    }
    Q4<T> dorn(axis.x, axis.y, axis.z, tcos(fAngle * dt * T(0.5)));
    Q4<T> p = qmul(dorn, q);
    T l = tsqrt(p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w);
    q = Q4<T>(p.x / l, p.y / l, p.z / l, p.w / l);
}
/* Single precision: the same map with the half angle h = |w| dt / 2 <= pi/8 (the clamp above) handled by even polynomials
   in h^2 for sin(h)/h and cos(h) (truncation < 1e-11), no square root or division on the common path, and the quaternion
   renormalised with the hardware reciprocal square root. */
TGB_D void integrateTransform(V3<float>& x, Q4<float>& q, const V3<float>& linvel, const V3<float>& angvel, float dt) {
    x = x + linvel * dt;
    const float w2 = angvel.length2();
    const float thr = SolverConst<float>::angularMotionThreshold();
    float h2 = w2 * (0.25f * dt * dt);         /* (|w| dt / 2)^2 */
    float scale = 0.5f * dt;                   /* axis = angvel * sin(h)/|w| = angvel * (dt/2) * sin(h)/h */
    if (w2 * dt * dt > thr * thr) {            /* rare: clamp the angle to pi/4 per step */
        const float fAngle = thr / dt;
        h2 = 0.25f * thr * thr;
        scale = 0.5f * dt * fAngle * rsqrtf(w2);
    }
    const float sinc = 1.0f + h2 * (-1.0f / 6.0f + h2 * (1.0f / 120.0f + h2 * (-1.0f / 5040.0f + h2 * (1.0f / 362880.0f))));
    const float cs = 1.0f + h2 * (-0.5f + h2 * (1.0f / 24.0f + h2 * (-1.0f / 720.0f + h2 * (1.0f / 40320.0f + h2 * (-1.0f / 3628800.0f)))));
    const V3<float> axis = angvel * (scale * sinc);
    Q4<float> p = qmul(Q4<float>(axis.x, axis.y, axis.z, cs), q);
    const float il = rsqrtf(p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w);
    q = Q4<float>(p.x * il, p.y * il, p.z * il, p.w * il);
}

/* btPlaneSpace1, first vector */
template <typename T>
TGB_D V3<T> planeSpace1(const V3<T>& n) {
    if (tabs(n.z) > T(0.7071067811865475244008443621048490)) {
        T a = n.y * n.y + n.z * n.z;
        T k = T(1) / tsqrt(a);
        return V3<T>(T(0), -n.z * k, n.y * k);
    }
    T a = n.x * n.x + n.y * n.y;
    T k = T(1) / tsqrt(a);
    return V3<T>(-n.y * k, n.x * k, T(0));
}

template <typename T> TGB_D T epsilonOf();
template <> TGB_D float epsilonOf<float>() { return 1.1920928955078125e-7f; }
template <> TGB_D double epsilonOf<double>() { return 2.2204460492503131e-16; }

/* sin of a double-precision phase: exact-type sin for double, range-reduced sinf for float */
TGB_D double sinPhase(double arg, double) { return sin(arg); }
TGB_D float sinPhase(double arg, float) {
    /* the phase grows with time: reduce it in double to [-pi, pi] before the single-precision sine */
    const double twoPi = 6.283185307179586476925286766559;
    const double r = arg - twoPi * rint(arg * (1.0 / twoPi));
    return sinf((float)r);
}

/* Projected Gauss-Seidel over contact rows kept in shared memory (bodies with more than two contact points: compound bodies,
   rods lying flat).  Inlined: a real call here costs more on the hot path (register shuffling around the call site; measured
   -5% on B200) than the cold code costs in instruction-cache footprint. */
template <typename T>
__device__ __forceinline__ void pgsSharedRows(T* ct, int rowStride, int nc, int iterations, bool anyPen, V3<T> gn, T invMass, T mu,
                                           V3<T>& dLin, V3<T>& dAng, V3<T>& push, V3<T>& turn) {
#define CTG(slot, f) ct[(slot) * rowStride + (f)]
    /* warm starting: the rows' initial impulses (CT_LAMN / CT_LAMT as the row set-up left them) act on the body straight away */
    for (int c = 0; c < nc; ++c) {
        const T lamN0 = CTG(c, CT_LAMN), lamT0 = CTG(c, CT_LAMT);
        const V3<T> t(CTG(c, CT_T + 0), CTG(c, CT_T + 1), CTG(c, CT_T + 2));
        const V3<T> angN(CTG(c, CT_ANGN + 0), CTG(c, CT_ANGN + 1), CTG(c, CT_ANGN + 2));
        const V3<T> angT(CTG(c, CT_ANGT + 0), CTG(c, CT_ANGT + 1), CTG(c, CT_ANGT + 2));
        dLin += gn * (invMass * lamN0) + t * (invMass * lamT0);
        dAng += angN * lamN0 + angT * lamT0;
    }
    if (anyPen)
        for (int it = 0; it < iterations; ++it) {
            for (int c = 0; c < nc; ++c) {
                const T rhsPen = CTG(c, CT_RHSPEN);
                if (rhsPen == T(0)) continue;
                const V3<T> rxn(CTG(c, CT_RXN + 0), CTG(c, CT_RXN + 1), CTG(c, CT_RXN + 2));
                const V3<T> angN(CTG(c, CT_ANGN + 0), CTG(c, CT_ANGN + 1), CTG(c, CT_ANGN + 2));
                const T invDiagN = CTG(c, CT_INVDIAGN);
                T delta = rhsPen;
                const T dv = gn.dot(push) + rxn.dot(turn);
                delta -= dv * invDiagN;
                const T lam = CTG(c, CT_LAMPUSH);
                const T sum = lam + delta;
                if (sum < T(0)) { delta = -lam; CTG(c, CT_LAMPUSH) = T(0); }
                else CTG(c, CT_LAMPUSH) = sum;
                push += gn * (invMass * delta);
                turn += angN * delta;
            }
        }
    for (int it = 0; it < iterations; ++it) {
        for (int c = 0; c < nc; ++c) {
            const V3<T> rxn(CTG(c, CT_RXN + 0), CTG(c, CT_RXN + 1), CTG(c, CT_RXN + 2));
            const V3<T> angN(CTG(c, CT_ANGN + 0), CTG(c, CT_ANGN + 1), CTG(c, CT_ANGN + 2));
            const T invDiagN = CTG(c, CT_INVDIAGN);
            T delta = CTG(c, CT_RHSN);
            const T dv = gn.dot(dLin) + rxn.dot(dAng);
            delta -= dv * invDiagN;
            const T lam = CTG(c, CT_LAMN);
            const T sum = lam + delta;
            if (sum < T(0)) { delta = -lam; CTG(c, CT_LAMN) = T(0); }
            else CTG(c, CT_LAMN) = sum;
            dLin += gn * (invMass * delta);
            dAng += angN * delta;
        }
        for (int c = 0; c < nc; ++c) {
            const T lamN = CTG(c, CT_LAMN);
            if (!(lamN > T(0))) continue;
            const V3<T> t(CTG(c, CT_T + 0), CTG(c, CT_T + 1), CTG(c, CT_T + 2));
            const V3<T> rxt(CTG(c, CT_RXT + 0), CTG(c, CT_RXT + 1), CTG(c, CT_RXT + 2));
            const V3<T> angT(CTG(c, CT_ANGT + 0), CTG(c, CT_ANGT + 1), CTG(c, CT_ANGT + 2));
            const T invDiagT = CTG(c, CT_INVDIAGT);
            const T lo = -mu * lamN, hi = mu * lamN;
            T delta = CTG(c, CT_RHST);
            const T dv = t.dot(dLin) + rxt.dot(dAng);
            delta -= dv * invDiagT;
            const T lam = CTG(c, CT_LAMT);
            const T sum = lam + delta;
            if (sum < lo) { delta = lo - lam; CTG(c, CT_LAMT) = lo; }
            else if (sum > hi) { delta = hi - lam; CTG(c, CT_LAMT) = hi; }
            else CTG(c, CT_LAMT) = sum;
            dLin += t * (invMass * delta);
            dAng += angT * delta;
        }
    }
#undef CTG
}

/* The same projected Gauss-Seidel for worlds with static obstacles: every row carries its own normal and friction coefficient
   (CT_NRM, CT_MU).  Row order = generation order: ground points first, then obstacle by obstacle. */
template <typename T>
__device__ __noinline__ void pgsRowsWithNormals(T* ct, int rowStride, int nc, int iterations, bool anyPen, T invMass,
                                                V3<T>& dLin, V3<T>& dAng, V3<T>& push, V3<T>& turn) {
#define CTG(slot, f) ct[(slot) * rowStride + (f)]
    /* warm starting: the rows' initial impulses act on the body straight away */
    for (int c = 0; c < nc; ++c) {
        const T lamN0 = CTG(c, CT_LAMN), lamT0 = CTG(c, CT_LAMT);
        const V3<T> n(CTG(c, CT_NRM + 0), CTG(c, CT_NRM + 1), CTG(c, CT_NRM + 2));
        const V3<T> t(CTG(c, CT_T + 0), CTG(c, CT_T + 1), CTG(c, CT_T + 2));
        const V3<T> angN(CTG(c, CT_ANGN + 0), CTG(c, CT_ANGN + 1), CTG(c, CT_ANGN + 2));
        const V3<T> angT(CTG(c, CT_ANGT + 0), CTG(c, CT_ANGT + 1), CTG(c, CT_ANGT + 2));
        dLin += n * (invMass * lamN0) + t * (invMass * lamT0);
        dAng += angN * lamN0 + angT * lamT0;
    }
    if (anyPen)
        for (int it = 0; it < iterations; ++it) {
            for (int c = 0; c < nc; ++c) {
                const T rhsPen = CTG(c, CT_RHSPEN);
                if (rhsPen == T(0)) continue;
                const V3<T> n(CTG(c, CT_NRM + 0), CTG(c, CT_NRM + 1), CTG(c, CT_NRM + 2));
                const V3<T> rxn(CTG(c, CT_RXN + 0), CTG(c, CT_RXN + 1), CTG(c, CT_RXN + 2));
                const V3<T> angN(CTG(c, CT_ANGN + 0), CTG(c, CT_ANGN + 1), CTG(c, CT_ANGN + 2));
                const T invDiagN = CTG(c, CT_INVDIAGN);
                T delta = rhsPen;
                const T dv = n.dot(push) + rxn.dot(turn);
                delta -= dv * invDiagN;
                const T lam = CTG(c, CT_LAMPUSH);
                const T sum = lam + delta;
                if (sum < T(0)) { delta = -lam; CTG(c, CT_LAMPUSH) = T(0); }
                else CTG(c, CT_LAMPUSH) = sum;
                push += n * (invMass * delta);
                turn += angN * delta;
            }
        }
    for (int it = 0; it < iterations; ++it) {
        for (int c = 0; c < nc; ++c) {
            const V3<T> n(CTG(c, CT_NRM + 0), CTG(c, CT_NRM + 1), CTG(c, CT_NRM + 2));
            const V3<T> rxn(CTG(c, CT_RXN + 0), CTG(c, CT_RXN + 1), CTG(c, CT_RXN + 2));
            const V3<T> angN(CTG(c, CT_ANGN + 0), CTG(c, CT_ANGN + 1), CTG(c, CT_ANGN + 2));
            const T invDiagN = CTG(c, CT_INVDIAGN);
            T delta = CTG(c, CT_RHSN);
            const T dv = n.dot(dLin) + rxn.dot(dAng);
            delta -= dv * invDiagN;
            const T lam = CTG(c, CT_LAMN);
            const T sum = lam + delta;
            if (sum < T(0)) { delta = -lam; CTG(c, CT_LAMN) = T(0); }
            else CTG(c, CT_LAMN) = sum;
            dLin += n * (invMass * delta);
            dAng += angN * delta;
        }
        for (int c = 0; c < nc; ++c) {
            const T lamN = CTG(c, CT_LAMN);
            if (!(lamN > T(0))) continue;
            const V3<T> t(CTG(c, CT_T + 0), CTG(c, CT_T + 1), CTG(c, CT_T + 2));
            const V3<T> rxt(CTG(c, CT_RXT + 0), CTG(c, CT_RXT + 1), CTG(c, CT_RXT + 2));
            const V3<T> angT(CTG(c, CT_ANGT + 0), CTG(c, CT_ANGT + 1), CTG(c, CT_ANGT + 2));
            const T invDiagT = CTG(c, CT_INVDIAGT);
            const T mu = CTG(c, CT_MU);
            const T lo = -mu * lamN, hi = mu * lamN;
            T delta = CTG(c, CT_RHST);
            const T dv = t.dot(dLin) + rxt.dot(dAng);
            delta -= dv * invDiagT;
            const T lam = CTG(c, CT_LAMT);
            const T sum = lam + delta;
            if (sum < lo) { delta = lo - lam; CTG(c, CT_LAMT) = lo; }
            else if (sum > hi) { delta = hi - lam; CTG(c, CT_LAMT) = hi; }
            else CTG(c, CT_LAMT) = sum;
            dLin += t * (invMass * delta);
            dAng += angT * delta;
        }
    }
#undef CTG
}

/* Signed distance of a point to an oriented box (centre co, axes e0 e1 e2, half extents h) and the outward direction at the closest
   surface point: outside, the vector from the closest point; inside, the axis of the nearest face (lowest index on a tie). */
template <typename T, bool kFast>
TGB_D T boxDistance(const V3<T>& p, const V3<T>& co, const V3<T>& e0, const V3<T>& e1, const V3<T>& e2, const V3<T>& h, V3<T>& nOut) {
    /* scalars and selects only: no indexed local arrays */
    const V3<T> d = p - co;
    const T q0 = e0.dot(d), q1 = e1.dot(d), q2 = e2.dot(d);
    const T a0 = tabs(q0) - h.x, a1 = tabs(q1) - h.y, a2 = tabs(q2) - h.z;
    const T o0 = a0 > T(0) ? a0 : T(0), o1 = a1 > T(0) ? a1 : T(0), o2 = a2 > T(0) ? a2 : T(0);
    const T len2 = (o0 * o0 + o1 * o1) + o2 * o2;
    T n0, n1, n2, sd;
    if (len2 > T(0)) {
        if (kFast) {
            /* single precision: one MUFU.RSQ instead of a square root and three divisions */
            const T inv = trsqrt(len2);
            sd = len2 * inv;
            n0 = (q0 < T(0) ? -o0 : o0) * inv;
            n1 = (q1 < T(0) ? -o1 : o1) * inv;
            n2 = (q2 < T(0) ? -o2 : o2) * inv;
        } else {
            sd = tsqrt(len2);
            n0 = (q0 < T(0) ? -o0 : o0) / sd;
            n1 = (q1 < T(0) ? -o1 : o1) / sd;
            n2 = (q2 < T(0) ? -o2 : o2) / sd;
        }
    } else {
        /* inside: the nearest face, lowest index on a tie */
        const bool k1 = a1 > a0;
        const T a01 = k1 ? a1 : a0;
        const bool k2 = a2 > a01;
        sd = k2 ? a2 : a01;
        n0 = (!k1 && !k2) ? (q0 < T(0) ? T(-1) : T(1)) : T(0);
        n1 = (k1 && !k2) ? (q1 < T(0) ? T(-1) : T(1)) : T(0);
        n2 = k2 ? (q2 < T(0) ? T(-1) : T(1)) : T(0);
    }
    nOut = e0 * n0 + e1 * n1 + e2 * n2;
    return sd;
}

/* Single precision, three or four contact points of one body (a rod lying flat on the ground, a vertebra on two rod ends): the same
   projected Gauss-Seidel in matrix form, everything in registers.  With A[r][s] = J_r M^-1 J_s^T over the rows (N0..N3, T0..T3) the
   relative velocity change dv_r of every row is carried along and updated by one multiply-add per applied impulse; a row visit is a
   dependent chain of ~5 operations instead of two dot products over shared-memory operands plus the body update (~15 operations and
   two shared-memory round trips).  Rows are visited in the order of pgsSharedRows (normals 0..nc-1, then frictions 0..nc-1): equal to
   it in exact arithmetic.  Measured need: warps holding a flat-lying rod ran 40 % longer than the median warp (TGB_STATS, round 2) --
   after the body pairs the second reason a launch waits for a few warps.  NORMALS: rows carry their own normal and friction
   coefficient (worlds with static obstacles).  Out of line: its registers are its own. */
template <typename T, bool NORMALS>
__device__ __noinline__ ContactSolveOut<T> solveContactsMatrix4(T* ct, int rowStride, int nc, int iterations, bool anyPen, V3<T> gn, T invMass, T muBody) {
    constexpr int M = 4, K = 8;
#define CTG(slot, f) ct[(slot) * rowStride + (f)]
    T A[K][K], rhs[K], invD[K], lam[K], rhsP[M], muv[M];
    {
        T jl[K][3], ja[K][3], bl[K][3], ba[K][3];
#pragma unroll
        for (int c = 0; c < M; ++c) {
            const bool has = c < nc;
            const int cc = has ? c : 0;
            const T z = has ? T(1) : T(0); /* rows this body does not have: zero Jacobian, nothing couples to them */
            const V3<T> n = NORMALS ? V3<T>(CTG(cc, CT_NRM + 0), CTG(cc, CT_NRM + 1), CTG(cc, CT_NRM + 2)) : gn;
            const T nn[3] = {n.x, n.y, n.z};
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                jl[c][j] = nn[j] * z; ja[c][j] = CTG(cc, CT_RXN + j) * z; ba[c][j] = CTG(cc, CT_ANGN + j) * z; bl[c][j] = jl[c][j] * invMass;
                jl[M + c][j] = CTG(cc, CT_T + j) * z; ja[M + c][j] = CTG(cc, CT_RXT + j) * z; ba[M + c][j] = CTG(cc, CT_ANGT + j) * z;
                bl[M + c][j] = jl[M + c][j] * invMass;
            }
            rhs[c] = has ? CTG(cc, CT_RHSN) : T(0); invD[c] = has ? CTG(cc, CT_INVDIAGN) : T(0); rhsP[c] = has ? CTG(cc, CT_RHSPEN) : T(0);
            rhs[M + c] = has ? CTG(cc, CT_RHST) : T(0); invD[M + c] = has ? CTG(cc, CT_INVDIAGT) : T(0);
            lam[c] = has ? CTG(cc, CT_LAMN) : T(0); lam[M + c] = has ? CTG(cc, CT_LAMT) : T(0); /* warm starting: what the row set-up left */
            muv[c] = NORMALS ? (has ? CTG(cc, CT_MU) : T(0)) : muBody;
        }
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
            for (int q = r; q < K; ++q) {
                const T v = (jl[r][0] * bl[q][0] + jl[r][1] * bl[q][1] + jl[r][2] * bl[q][2]) + (ja[r][0] * ba[q][0] + ja[r][1] * ba[q][1] + ja[r][2] * ba[q][2]);
                A[r][q] = v;
                A[q][r] = v;
            }
    }
    V3<T> dLin(T(0), T(0), T(0)), dAng(T(0), T(0), T(0)), push(T(0), T(0), T(0)), turn(T(0), T(0), T(0));
    if (anyPen) {
        /* split impulse: the normal rows alone, with their own impulses and relative velocities */
        T lamP[M], dvP[M];
#pragma unroll
        for (int c = 0; c < M; ++c) { lamP[c] = T(0); dvP[c] = T(0); }
        for (int it = 0; it < iterations; ++it) {
#pragma unroll
            for (int c = 0; c < M; ++c) {
                if (c >= nc || rhsP[c] == T(0)) continue;
                T delta = rhsP[c] - dvP[c] * invD[c];
                const T sum = lamP[c] + delta;
                if (sum < T(0)) { delta = -lamP[c]; lamP[c] = T(0); }
                else lamP[c] = sum;
#pragma unroll
                for (int r = 0; r < M; ++r) dvP[r] += A[r][c] * delta;
            }
        }
#pragma unroll
        for (int c = 0; c < M; ++c) {
            if (c >= nc) continue;
            const V3<T> n = NORMALS ? V3<T>(CTG(c, CT_NRM + 0), CTG(c, CT_NRM + 1), CTG(c, CT_NRM + 2)) : gn;
            push += n * (invMass * lamP[c]);
            turn += V3<T>(CTG(c, CT_ANGN + 0), CTG(c, CT_ANGN + 1), CTG(c, CT_ANGN + 2)) * lamP[c];
        }
    }
    T dv[K];
#pragma unroll
    for (int r = 0; r < K; ++r) {
        T v = T(0);
#pragma unroll
        for (int q = 0; q < K; ++q) v += A[r][q] * lam[q];
        dv[r] = v;
    }
    for (int it = 0; it < iterations; ++it) {
#pragma unroll
        for (int c = 0; c < M; ++c) {
            if (c >= nc) continue;
            T delta = rhs[c] - dv[c] * invD[c];
            const T sum = lam[c] + delta;
            if (sum < T(0)) { delta = -lam[c]; lam[c] = T(0); }
            else lam[c] = sum;
#pragma unroll
            for (int r = 0; r < K; ++r) dv[r] += A[r][c] * delta;
        }
#pragma unroll
        for (int c = 0; c < M; ++c) {
            if (c >= nc || !(lam[c] > T(0))) continue;
            const T lo = -muv[c] * lam[c], hi = muv[c] * lam[c];
            T delta = rhs[M + c] - dv[M + c] * invD[M + c];
            const T sum = lam[M + c] + delta;
            if (sum < lo) { delta = lo - lam[M + c]; lam[M + c] = lo; }
            else if (sum > hi) { delta = hi - lam[M + c]; lam[M + c] = hi; }
            else lam[M + c] = sum;
#pragma unroll
            for (int r = 0; r < K; ++r) dv[r] += A[r][M + c] * delta;
        }
    }
    /* the body's velocity change from the accumulated impulses; the final impulses go back into the rows (persistent manifold) */
#pragma unroll
    for (int c = 0; c < M; ++c) {
        if (c >= nc) continue;
        const V3<T> n = NORMALS ? V3<T>(CTG(c, CT_NRM + 0), CTG(c, CT_NRM + 1), CTG(c, CT_NRM + 2)) : gn;
        const V3<T> t(CTG(c, CT_T + 0), CTG(c, CT_T + 1), CTG(c, CT_T + 2));
        dLin += n * (invMass * lam[c]) + t * (invMass * lam[M + c]);
        dAng += V3<T>(CTG(c, CT_ANGN + 0), CTG(c, CT_ANGN + 1), CTG(c, CT_ANGN + 2)) * lam[c] +
                V3<T>(CTG(c, CT_ANGT + 0), CTG(c, CT_ANGT + 1), CTG(c, CT_ANGT + 2)) * lam[M + c];
        CTG(c, CT_LAMN) = lam[c]; CTG(c, CT_LAMT) = lam[M + c];
    }
#undef CTG
    ContactSolveOut<T> o;
    o.dLin = dLin; o.dAng = dAng; o.push = push; o.turn = turn;
    return o;
}

/* Contact solve of a body with two or more contact points (a rod lying on a rim line, a compound body): rare next to the
   single-point case, so it lives out of line -- the step loop's hot instructions stay within the instruction cache.
   ct = row 0 of this body, rowStride = words between consecutive rows of one body. */
template <typename T, bool kFast>
__device__ __noinline__ ContactSolveOut<T> solveContactsGeneral(T* ct, int rowStride, int nc, int iterations, bool anyPen, V3<T> gn, T invMass,
                                                                 T mu) {
    V3<T> dLin(T(0), T(0), T(0)), dAng(T(0), T(0), T(0)), push(T(0), T(0), T(0)), turn(T(0), T(0), T(0));
    if (nc <= 2) {
        /* one or two contact points (a rod touching with one rim point or lying on a rim line): the rows are read into
           registers once and all iterations run there; same operations in the same order as the general path below */
        T rxn[2][3], angN[2][3], invDiagN[2], rhsN[2], rhsPen[2], lamN[2], lamPush[2];
        T tt[2][3], rxt[2][3], angT[2][3], invDiagT[2], rhsT[2], lamT[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int cc = c < nc ? c : 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                rxn[c][j] = ct[(cc) * rowStride + (CT_RXN + j)]; angN[c][j] = ct[(cc) * rowStride + (CT_ANGN + j)];
                tt[c][j] = ct[(cc) * rowStride + (CT_T + j)]; rxt[c][j] = ct[(cc) * rowStride + (CT_RXT + j)]; angT[c][j] = ct[(cc) * rowStride + (CT_ANGT + j)];
            }
            invDiagN[c] = ct[(cc) * rowStride + (CT_INVDIAGN)]; rhsN[c] = ct[(cc) * rowStride + (CT_RHSN)]; rhsPen[c] = ct[(cc) * rowStride + (CT_RHSPEN)];
            invDiagT[c] = ct[(cc) * rowStride + (CT_INVDIAGT)]; rhsT[c] = ct[(cc) * rowStride + (CT_RHST)];
            /* warm starting: the initial impulses the row set-up left in the row */
            lamN[c] = c < nc ? ct[(cc) * rowStride + (CT_LAMN)] : T(0); lamPush[c] = T(0); lamT[c] = c < nc ? ct[(cc) * rowStride + (CT_LAMT)] : T(0);
        }
        if (anyPen)
            for (int it = 0; it < iterations; ++it) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c >= nc || rhsPen[c] == T(0)) continue;
                    T delta = rhsPen[c];
                    const T dv = gn.dot(push) + (rxn[c][0] * turn.x + rxn[c][1] * turn.y + rxn[c][2] * turn.z);
                    delta -= dv * invDiagN[c];
                    const T sum = lamPush[c] + delta;
                    if (sum < T(0)) { delta = -lamPush[c]; lamPush[c] = T(0); }
                    else lamPush[c] = sum;
                    push += gn * (invMass * delta);
                    turn += V3<T>(angN[c][0], angN[c][1], angN[c][2]) * delta;
                }
            }
        if (kFast && TGB_MATRIX_PGS) {
            /* Single precision: the same projected Gauss-Seidel in matrix form.  With A[r][s] = J_r M^-1 J_s^T over the
               rows (N0, N1, T0, T1) the relative velocity change dv_r of every row is carried along and updated by one
               multiply-add per applied impulse, instead of being re-formed from the body's velocity change by two dot
               products: the dependent chain per row drops from ~10 to ~5 operations and the instruction count by ~4x.
               Equal in exact arithmetic to the sequential-impulse form (kept for double precision, the parity path). */
            T jl[4][3], ja[4][3], bl[4][3], ba[4][3], A[4][4], dv[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                jl[c][0] = gn.x; jl[c][1] = gn.y; jl[c][2] = gn.z;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    ja[c][j] = rxn[c][j]; ba[c][j] = angN[c][j]; bl[c][j] = jl[c][j] * invMass;
                    jl[2 + c][j] = tt[c][j]; ja[2 + c][j] = rxt[c][j]; ba[2 + c][j] = angT[c][j]; bl[2 + c][j] = tt[c][j] * invMass;
                }
            }
            const bool two = nc > 1;
            /* A is symmetric: the upper triangle is computed (only the rows of the first contact when there is one) */
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = r; q < 4; ++q) {
                    T v = T(0);
                    if (two || ((r & 1) == 0 && (q & 1) == 0))
                        v = (jl[r][0] * bl[q][0] + jl[r][1] * bl[q][1] + jl[r][2] * bl[q][2]) +
                            (ja[r][0] * ba[q][0] + ja[r][1] * ba[q][1] + ja[r][2] * ba[q][2]);
                    A[r][q] = v;
                    A[q][r] = v;
                }
            /* warm starting: the rows' relative velocity changes under the initial impulses (rows N0, N1, T0, T1) */
#pragma unroll
            for (int r = 0; r < 4; ++r) dv[r] = (A[r][0] * lamN[0] + A[r][1] * lamN[1]) + (A[r][2] * lamT[0] + A[r][3] * lamT[1]);
            for (int it = 0; it < iterations; ++it) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c == 1 && !two) continue;
                    T delta = rhsN[c] - dv[c] * invDiagN[c];
                    const T sum = lamN[c] + delta;
                    if (sum < T(0)) { delta = -lamN[c]; lamN[c] = T(0); }
                    else lamN[c] = sum;
#pragma unroll
                    for (int r = 0; r < 4; ++r) dv[r] += A[r][c] * delta;
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if ((c == 1 && !two) || !(lamN[c] > T(0))) continue;
                    const T lo = -mu * lamN[c], hi = mu * lamN[c];
                    T delta = rhsT[c] - dv[2 + c] * invDiagT[c];
                    const T sum = lamT[c] + delta;
                    if (sum < lo) { delta = lo - lamT[c]; lamT[c] = lo; }
                    else if (sum > hi) { delta = hi - lamT[c]; lamT[c] = hi; }
                    else lamT[c] = sum;
#pragma unroll
                    for (int r = 0; r < 4; ++r) dv[r] += A[r][2 + c] * delta;
                }
            }
            /* the body's velocity change from the accumulated impulses */
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (c == 1 && !two) continue;
                dLin += V3<T>(bl[c][0], bl[c][1], bl[c][2]) * lamN[c] + V3<T>(bl[2 + c][0], bl[2 + c][1], bl[2 + c][2]) * lamT[c];
                dAng += V3<T>(ba[c][0], ba[c][1], ba[c][2]) * lamN[c] + V3<T>(ba[2 + c][0], ba[2 + c][1], ba[2 + c][2]) * lamT[c];
            }
        } else {
        /* warm starting: the initial impulses act on the body straight away (contact by contact, the oracle's order) */
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (c >= nc) continue;
            dLin += gn * (invMass * lamN[c]) + V3<T>(tt[c][0], tt[c][1], tt[c][2]) * (invMass * lamT[c]);
            dAng += V3<T>(angN[c][0], angN[c][1], angN[c][2]) * lamN[c] + V3<T>(angT[c][0], angT[c][1], angT[c][2]) * lamT[c];
        }
        for (int it = 0; it < iterations; ++it) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (c >= nc) continue;
                T delta = rhsN[c];
                const T dv = gn.dot(dLin) + (rxn[c][0] * dAng.x + rxn[c][1] * dAng.y + rxn[c][2] * dAng.z);
                delta -= dv * invDiagN[c];
                const T sum = lamN[c] + delta;
                if (sum < T(0)) { delta = -lamN[c]; lamN[c] = T(0); }
                else lamN[c] = sum;
                dLin += gn * (invMass * delta);
                dAng += V3<T>(angN[c][0], angN[c][1], angN[c][2]) * delta;
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (c >= nc || !(lamN[c] > T(0))) continue;
                const V3<T> t(tt[c][0], tt[c][1], tt[c][2]);
                const T lo = -mu * lamN[c], hi = mu * lamN[c];
                T delta = rhsT[c];
                const T dv = t.dot(dLin) + (rxt[c][0] * dAng.x + rxt[c][1] * dAng.y + rxt[c][2] * dAng.z);
                delta -= dv * invDiagT[c];
                const T sum = lamT[c] + delta;
                if (sum < lo) { delta = lo - lamT[c]; lamT[c] = lo; }
                else if (sum > hi) { delta = hi - lamT[c]; lamT[c] = hi; }
                else lamT[c] = sum;
                dLin += t * (invMass * delta);
                dAng += V3<T>(angT[c][0], angT[c][1], angT[c][2]) * delta;
            }
        }
        }
        /* the final impulses go back into the rows: the persistent manifold keeps them for the next step */
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (c < nc) { ct[c * rowStride + CT_LAMN] = lamN[c]; ct[c * rowStride + CT_LAMT] = lamT[c]; }
    } else if (kFast && TGB_MATRIX_PGS && TGB_MATRIX4 && nc <= 4) {
        return solveContactsMatrix4<T, false>(ct, rowStride, nc, iterations, anyPen, gn, invMass, mu);
    } else {
        pgsSharedRows<T>(ct, rowStride, nc, iterations, anyPen, gn, invMass, mu, dLin, dAng, push, turn);
    }
    ContactSolveOut<T> o;
    o.dLin = dLin; o.dAng = dAng; o.push = push; o.turn = turn;
    return o;
}

/* One cable's working set in registers (the cable phase keeps several of these in flight per lane). */
template <typename T>
struct CableRegs {
    int c, actuator, slots, ia, ib;
    bool valid;
    V3<T> ra, rb, dist, J;
    T L, k, damp, tVel, maxTens, maxStretch, minAL, minRL;
    T rest, prev, pref, motorVel, desiredTorque, velocity;
    /* tgBasicActuator::moveMotors, tgBasicActuator.cpp:170-225 */
    TGB_D void moveMotors(T dtm, bool enable = true, T stepScale = T(1)) {
        /* written with selects only: the same values as the reference's branches, no divergence in the cable phase.
           stepScale = 2 stands for two consecutive calls with the same preferred length (EscapeController moves every motor twice
           per step): two clamped moves towards one target are one move clamped at twice the step -- the tension clamp of the
           preferred length is idempotent, the sign of the remaining difference and the length test do not change between the
           calls, and the minimum rest length clamps the sum the same way.  Equal in exact arithmetic; used in single precision. */
        const T stepSize = tVel * dtm * stepScale;
        const T newPref = ((L - pref) * k > maxTens) ? L - maxStretch : pref;
        const T diff = newPref - rest;
        const T move = (tabs(diff) > stepSize) ? tcopysign(stepSize, diff) : diff; /* (diff / |diff|) * stepSize, exactly */
        T newRest = ((L > minAL) || (diff > T(0))) ? rest + move : rest;
        newRest = (newRest > minRL) ? newRest : minRL;
        pref = enable ? newPref : pref;
        rest = enable ? newRest : rest;
    }
};

/* CPG coupling lists of up to 512 words are staged into shared memory (see the kernel) */
__host__ __device__ inline bool tgb_cpg_topo_staged(const tgb_step_args& a) {
    return a.ctrl_kind == TGB_CTRL_CPG_IMPEDANCE && (a.C + 1) + 2 * a.cpg_n_couplings <= 512;
}

/* doubles of shared memory per env group: controller row, sine cumulative phases, CPG work space */
__host__ __device__ inline int tgb_group_double_words(const tgb_step_args& a, int lpe) {
    int n = a.ctrl_stride + 8;
    if (a.ctrl_kind == TGB_CTRL_ESCAPE_SINE) n += TGB_SINE_REC * 24 + 2 * 8; /* per cable (sin, cos[, amplitude, offset]) of the phase, per cluster (cos, sin) of the phase step */
    if (a.ctrl_kind == TGB_CTRL_CPG_IMPEDANCE) n += 3 * a.cpg_n_pairs + (3 + 3 + 3 + 21 + 1 + 3) * a.C;
    n = (n + 1) & ~1; /* even: the int region behind the doubles stays 16-byte aligned */
    /* A 64-bit access of a warp goes through the banks half a warp at a time; lanes of a group read consecutive doubles (the sine
       controller's phase pairs), so the groups that share a half warp must start lpe doubles apart modulo the 16 doubles of a bank
       row (ncu, rounds 1 and 2: LDS.64 / STS.64 of the phase pairs at 50 % excessive wavefronts with the regions 176 doubles apart) */
    if (lpe < 16 && a.ctrl_kind == TGB_CTRL_ESCAPE_SINE) while ((n & 15) != lpe) n += 2;
    return n;
}

/* Views of the packed model in shared memory. */
template <typename T>
struct ModelView {
    const T* bt; const T* st; const T* ct;
    const int* bi; const int* ci; const int* inc;
    const T* crec; const int* cirec; /* per-cable records of the cable phase */
    int R, S, C;
    TGB_D T body(int f, int b) const { return bt[f * R + b]; }
    TGB_D int bodyi(int f, int b) const { return bi[f * R + b]; }
    TGB_D T shape(int f, int s) const { return st[f * S + s]; }
    TGB_D T cable(int f, int c) const { return ct[f * C + c]; }
    TGB_D int cablei(int f, int c) const { return ci[f * C + c]; }
};

/* Constants of one collision shape (a cylinder in the body frame). */
template <typename T>
struct ShapeConst {
    T radius, halfLen, margin;
    V3<T> spos, r0, r1; /* position, columns 0 and 1 of the rotation */
    bool sphere;        /* a point with margin = radius: one candidate contact point instead of eight rim points */
    bool box;           /* half extents (radius, halfLen, halfZ): the eight corners of the core box instead of rim points */
    T halfZ;
};

/* Controller classes a kernel is compiled for (the controller kind inside a class is a warp-uniform run-time switch):
   SIMPLE = none / set-point / torque / tension, SINE = EscapeController, CPG = CPG + impedance (its Dormand-Prince integrator
   needs 231 registers).  GROUND = false compiles the contact manifold and the contact solver out (tgEmptyGround).  Keeping the
   hot loop's code small matters: it is ~20 KB of instructions against a 32 KB L1.5 instruction cache, and at 4096 environments
   there are too few warps per SM to hide instruction-fetch misses. */
enum { TGB_CLASS_SIMPLE = 0, TGB_CLASS_SINE = 1, TGB_CLASS_CPG = 2 };
TGB_HOST_DEVICE_INLINE int tgb_ctrl_class(int kind) {
    return kind == TGB_CTRL_CPG_IMPEDANCE ? TGB_CLASS_CPG : kind == TGB_CTRL_ESCAPE_SINE ? TGB_CLASS_SINE : TGB_CLASS_SIMPLE;
}

/* WIDE = true is the layout for few environments (the machine is latency bound: ~1.7 warps per scheduler at 4096 environments): a
   CTA of four warps holds 32 / LPE environments; warp 0 is the body warp (lane j of a group of LPE = body j, as in the narrow
   layout) and all four warps share the cable phase with 4 LPE lanes per environment, so that the cables of an environment run in
   one round instead of C / LPE rounds.  The two per-step warp barriers become CTA barriers.  The arithmetic of a body / cable does
   not depend on the lane that runs it: results are bit-identical to the narrow layout. */
/* Register budgets: the narrow layout is capped by shared memory at 8 one-warp CTAs per SM, so its lanes may take what ptxas wants
   (minBlocks 1: 209 registers instead of 163, +3.3 % on B200); the four-warp layout needs three CTAs per SM (168 registers). */
#ifndef TGB_FUSE_DOUBLE_MOVE
#define TGB_FUSE_DOUBLE_MOVE 1
#endif
#ifndef TGB_WIDE_MINBLOCKS
#define TGB_WIDE_MINBLOCKS 3
#endif
#ifndef TGB_NARROW_MINBLOCKS
#define TGB_NARROW_MINBLOCKS 1
#endif
template <typename T, int LPE, int CLASS, bool GROUND, int WIDE = 0 /* 0 or 4 warps */, bool OBST = false /* static obstacles, narrow layout only */,
          bool WARM = false /* persistent manifolds with warm starting (tgb_set_warm_starting), narrow layout only: compiled out of the default
                               kernels -- even skipped at run time its code cost 6 % of the step (registers, instruction footprint) */>
__global__ void __launch_bounds__(128, WIDE ? TGB_WIDE_MINBLOCKS : TGB_NARROW_MINBLOCKS) tgb_step_kernel(const tgb_step_args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef typename Vec16<T>::type vec_t;
    const int VN = Vec16<T>::N;
    constexpr bool kFast = std::is_same<T, float>::value; /* float: reassociated sums and fast reciprocals; double: the oracle's order */

    const int tid = threadIdx.x;
    const int lane = tid % LPE;
    const bool bodyWarp = !WIDE || tid < 32; /* WIDE: warps 1..3 shadow the groups of warp 0 in the body role and never store */
    const int grp = WIDE ? (tid & 31) / LPE : tid / LPE;
    const int groupsPerBlock = WIDE ? 32 / LPE : blockDim.x / LPE;
    /* launches over a list of environments: slot -> environment through env_list; CTAs beyond the list's length leave at once
       (before any barrier), groups beyond it shadow the list's last environment */
    int listCount = 0;
    if (a.env_list != nullptr) {
        listCount = *a.env_list_count;
        if (listCount > a.env_list_slots) listCount = a.env_list_slots;
        if ((int)blockIdx.x * groupsPerBlock >= listCount) return;
    }
    const int slotRaw = blockIdx.x * groupsPerBlock + grp;
    const int envRaw = a.env_list != nullptr ? (slotRaw < listCount ? a.env_list[slotRaw] : a.n_env) : a.env_begin + slotRaw;
    const int envShadow = a.env_list != nullptr ? a.env_list[listCount - 1] : a.n_env - 1;
    const bool live = bodyWarp && envRaw < a.n_env;
    const int env = envRaw < a.n_env ? envRaw : envShadow; /* tail groups shadow the last env and never store */
    /* cable role: the lanes that share an environment's cables */
    constexpr int CL = WIDE ? WIDE * LPE : LPE;
    const int cLane = WIDE ? tid % CL : lane;
    const int cGrp = WIDE ? tid / CL : grp;
    const int cSlotRaw = blockIdx.x * groupsPerBlock + cGrp;
    const int cEnvRaw = a.env_list != nullptr ? (cSlotRaw < listCount ? a.env_list[cSlotRaw] : a.n_env) : a.env_begin + cSlotRaw;
    const bool cLive = cEnvRaw < a.n_env;
    const int cEnv = cLive ? cEnvRaw : envShadow;
    /* lanes of this group inside the warp, for shuffles */
    const unsigned fullMask = 0xffffffffu;
    const int warpLane = tid & 31;
    const int groupLane0 = warpLane - lane;

    const int R = a.R, C = a.C, S = a.S;

    /* ---- shared memory carve-up ----------------------------------------------------------------------- */
    /* [model T][per-group T regions][per-group double regions][model ints] */
    T* sModelT = reinterpret_cast<T*>(smem_raw);
    const int modelTWords = (a.model_t_words + 3) & ~3;
    const int poseWords = 12 * LPE;
    /* impulse exchange between the cable phase and the bodies.  Double: J per cable [C][3], gathered through the incidence list.
       Float: every cable end writes (J, torque about the body's centre) into the slot [position in the body's incidence
       list][body lane] of two float4 planes, so that the gather of a body lane is a run of 128-bit loads at consecutive addresses
       across lanes: no index loads, no bank conflicts. */
    const int jWords = kFast ? a.max_inc * LPE * 8 : 3 * C;
    /* without the persistent manifolds a row has no key: 23 words (the stride must stay odd, see CT_*) */
    constexpr int NCT = OBST ? (int)TGB_NCT_OBST : (WARM ? (int)TGB_NCT : (int)TGB_NCT - 2);
    const int rowBodies = OBST ? a.n_dyn : R;     /* worlds with obstacles: no rows for the static bodies */
    const int ctOnlyWords = a.cap_contacts * NCT * rowBodies; /* rows [slot][body][NCT]: 23 words apart across lanes, conflict-free */
    const int pairWords = (a.n_pairs > 0 && a.pair_cap > 0) ? a.pair_cap * (int)TGB_NPR : 0; /* rows between two bodies, behind the contact rows */
    const int ctWords = ctOnlyWords + pairWords;
    /* the contact rows (live from contact generation to the end of the solver) and the cable impulses (live from the cable
       phase to the impulse gather) share one region: shared memory per environment decides how many CTAs an SM holds */
    const int exchangeWords = TGB_EXCHANGE_UNION ? (jWords > ctWords ? jWords : ctWords) : jWords + ctWords;
    /* body pairs: per pair the slack left before it has to be tested again, per body the motion bound of the last step (tgb_pairs.cuh) */
    const int pairKeepWords = a.n_pairs > 0 ? a.n_pairs + LPE : 0; /* slack per pair, motion bound per body */
    /* persistent manifolds (warm starting): the rows' impulses of the last world step, see TGB_WARM_* in tgb_dev.h */
#define warmOn (WARM && a.warm != nullptr) /* the host passes the buffer only with a positive factor */
#define warmPairCap ((a.n_pairs > 0 && a.pair_cap > 0) ? a.pair_cap : 0)
#define warmWords (warmOn ? TGB_WARM_STATE_WORDS(a.cap_contacts, rowBodies, warmPairCap) : 0)
    int groupTWords = a.stride + poseWords + exchangeWords + pairKeepWords + warmWords;
    groupTWords = (groupTWords + 3) & ~3;
    /* keep group regions LPE banks apart so that the groups of one warp do not collide */
    if (LPE < 32) { while ((groupTWords & 31) != (LPE < 4 ? 4 : LPE)) groupTWords += 4; }
    T* sGroupT = sModelT + modelTWords + grp * groupTWords;
    T* sState = sGroupT;                       /* env record: bodies then cables */
    T* sPose = sState + a.stride;              /* [LPE][12] : pos3, basis9 of each body, one 16-byte aligned record per body */
    T* sJ = sPose + poseWords;                 /* [C][3] */
    T* sCt = TGB_EXCHANGE_UNION ? sJ : sJ + jWords; /* [cap][R][TGB_NCT], same region */
    T* sPairRows = sCt + ctOnlyWords;               /* [pair_cap][TGB_NPR] */
    T* sPairSlack = sJ + exchangeWords;             /* [n_pairs], then [LPE] motion bounds */
    T* sMotion = sPairSlack + a.n_pairs;
#define sWarm (sPairSlack + pairKeepWords)         /* [cap_contacts][row bodies][3], then [pair_cap][3] */
#define sWarmPair (sWarm + TGB_WARM_WORDS * a.cap_contacts * rowBodies)
    constexpr bool cpgOn = CLASS == TGB_CLASS_CPG;
    constexpr bool sineOn = CLASS == TGB_CLASS_SINE;
    constexpr bool simpleOn = CLASS == TGB_CLASS_SIMPLE;
    /* cables in flight per lane in the cable phase: 3 where the instruction footprint allows it (no contact code), else 1 --
       measured on B200: 3 is +7% without ground and -10% with ground (the hot loop then overflows the instruction cache) */
    constexpr int CU = TGB_CABLE_UNROLL ? TGB_CABLE_UNROLL : (GROUND ? 1 : 3);
    const int groupDWords = tgb_group_double_words(a, LPE);
    double* sD = reinterpret_cast<double*>(sModelT + modelTWords + groupsPerBlock * groupTWords) + grp * groupDWords;
    /* the same regions of the cable role's environment */
    T* cStateP = sModelT + modelTWords + cGrp * groupTWords;
    T* cPoseP = cStateP + a.stride;
    T* cJP = cPoseP + poseWords;
    double* cCtrlP = reinterpret_cast<double*>(sModelT + modelTWords + groupsPerBlock * groupTWords) + cGrp * groupDWords;
    double* cCumP = cCtrlP + a.ctrl_stride;
    double* cSineP = cCumP + 8;
    double* cRotP = cSineP + TGB_SINE_REC * 24;
    double* sCtrl = sD;                        /* controller parameter row */
    double* sCum = sD + a.ctrl_stride;         /* sine controller: cumulative phase per cluster */
    /* CPG: pair table (weight, cos phase, sin phase), state x, stage state, published (sin, cos, r), k1..k7, node values */
    /* sine controller, single precision: (sin, cos) of each cable's phase advanced by a rotation per step, see below */
    double* sSine = sCum + 8;                  /* [TGB_SINE_REC][24] planes: sin, cos[, amplitude, offset] (consecutive cables on consecutive banks) */
    double* sRot = sSine + TGB_SINE_REC * 24;             /* [8][2] cos, sin of frequency * dt */
    double* sPair = sCum + 8;
    double* sX = sPair + 3 * a.cpg_n_pairs;
    double* sXt = sX + 3 * C;
    double* sTrig = sXt + 3 * C;
    double* sK = sTrig + 3 * C;                /* [7][3 C] */
    double* sNodeValue = sK + 21 * C;
    double* sPid = sNodeValue + C;             /* [C][3] tgPIDController previous error, integrated error; last commanded tension */
    int* sModelI = reinterpret_cast<int*>(reinterpret_cast<double*>(sModelT + modelTWords + groupsPerBlock * groupTWords) +
                                          groupsPerBlock * groupDWords);

    /* ---- stage the model and this group's env record ----------------------------------------------------- */
    {
        const T* gmt = reinterpret_cast<const T*>(a.model_t);
        for (int i = tid; i < a.model_t_words; i += blockDim.x) sModelT[i] = gmt[i];
        for (int i = tid; i < a.model_i_words; i += blockDim.x) sModelI[i] = a.model_i[i];
        const vec_t* gs = reinterpret_cast<const vec_t*>(reinterpret_cast<const T*>(a.state) + (size_t)env * a.stride);
        vec_t* ss = reinterpret_cast<vec_t*>(sState);
        for (int i = lane; i < a.stride / VN; i += LPE) ss[i] = gs[i];
        if (a.ctrl != nullptr)
            for (int i = lane; i < a.ctrl_stride; i += LPE) sCtrl[i] = a.ctrl[(size_t)env * a.ctrl_stride + i];
        if (warmOn && bodyWarp) {
            const T* gw = reinterpret_cast<const T*>(a.warm) + (size_t)env * warmWords;
            for (int i = lane; i < warmWords; i += LPE) sWarm[i] = gw[i];
        }
    }
    __syncthreads();

    ModelView<T> M;
    M.R = R; M.S = S; M.C = C;
    M.bt = sModelT; M.st = M.bt + TGB_NBF * R; M.ct = M.st + TGB_NSF * S;
    M.bi = sModelI; M.ci = M.bi + TGB_NBI * R; M.inc = M.ci + TGB_NCI * C;
    M.crec = sModelT + a.model_crec_off; M.cirec = sModelI + a.model_cirec_off;

    const T dt = (T)a.dt;
    const T invDt = (T)(1.0 / a.dt);
    const V3<T> gn((T)a.gn[0], (T)a.gn[1], (T)a.gn[2]);
    const int ctrlKind = a.ctrl_kind;

    /* env frame (every lane of the group keeps an identical copy) */
    double Ox, Oy, Oz, Vx, Vy, Vz, ctrlTime, frameSpare;
    {
        const double* f = a.frame + (size_t)env * TGB_NFR;
        frameSpare = f[FR_SPARE];
        Ox = f[FR_O]; Oy = f[FR_O + 1]; Oz = f[FR_O + 2];
        Vx = f[FR_V]; Vy = f[FR_V + 1]; Vz = f[FR_V + 2];
        ctrlTime = f[FR_CTRLTIME];
    }

    if (sineOn && lane == 0) {
        /* EscapeController.cpp:279-290: the phase grows by the cluster's phase change at every cluster boundary */
        double ph = 0.0;
        sCum[0] = 0.0;
        for (int c = 1; c < 8; ++c) { ph += sCtrl[c * 4 + 2]; sCum[c] = ph; }
    }
    /* Single precision: the sine of each cable's phase is not re-evaluated every step but advanced by the rotation
       (cos, sin)(frequency dt) in double precision (error ~1e-16 per step, far below single-precision resolution; saves a
       double-precision range reduction and a sine per cable and step).  The (sin, cos) pairs start from exact values when the
       controller is installed and travel through HBM between launches, so results do not depend on how an episode is cut
       into launches. */
    const bool sineRecurrence = sineOn && kFast && a.sine_state != nullptr;
    if (sineRecurrence) {
        __syncwarp();
        double* gsine = a.sine_state + (size_t)env * 48;
        for (int c = lane; c < 24 && c < C; c += LPE) {
            const int cluster = c < 16 ? (c >> 2) : 4 + ((c - 16) >> 1);
            double sn, cn;
            if (frameSpare != 0.0) {
                sn = gsine[2 * c]; cn = gsine[2 * c + 1];
            } else {
                const bool firstOfCluster = c < 16 ? ((c & 3) == 0) : (((c - 16) & 1) == 0);
                const double phase = (firstOfCluster && cluster > 0) ? sCum[cluster - 1] : sCum[cluster];
                sincos(sCtrl[cluster * 4 + 1] * ctrlTime + phase, &sn, &cn);
            }
            sSine[c] = sn; sSine[24 + c] = cn;
            if (TGB_SINE_REC == 4) { sSine[48 + c] = sCtrl[cluster * 4 + 0]; sSine[72 + c] = sCtrl[cluster * 4 + 3]; }
        }
        for (int cl = lane; cl < 8; cl += LPE) {
            double sn, cn;
            sincos(sCtrl[cl * 4 + 1] * a.dt, &sn, &cn);
            sRot[2 * cl] = cn; sRot[2 * cl + 1] = sn;
        }
        __syncwarp();
    }

    double cpgUpdateTime = 0.0, cpgCableTime = 0.0;
    if (cpgOn) {
        const double* cs = a.cpg_state + (size_t)env * TGB_CPG_STATE_DOUBLES(C);
        for (int i = lane; i < 3 * C; i += LPE) sPid[i] = cs[3 * C + 2 + i];
        for (int i = lane; i < 3 * C; i += LPE) sX[(i % 3) * C + i / 3] = cs[i]; /* node-major in HBM, planes [variable][node] here */
        cpgUpdateTime = cs[3 * C];
        cpgCableTime = cs[3 * C + 1];
        for (int p = lane; p < a.cpg_n_pairs; p += LPE) {
            double sn, cn;
            sincos(sCtrl[2 + 2 * p + 1], &sn, &cn);
            sPair[p] = sCtrl[2 + 2 * p];             /* planes [weight | cos | sin][pair]: lanes read consecutive pairs */
            sPair[a.cpg_n_pairs + p] = cn;
            sPair[2 * a.cpg_n_pairs + p] = sn;
        }
        __syncwarp();
        for (int i = lane; i < C; i += LPE) sNodeValue[i] = sX[C + i] * cos(sX[i]); /* CPGNode::updateNodeValues */
    }
    int statusAcc = 0;

    /* ---- body state into registers ------------------------------------------------------------------------ */
    const bool hasBody = bodyWarp && lane < R;
    const int bidx = hasBody ? lane : 0;
    BodyReg<T> b;
    {
        const T* s = sState + bidx * TGB_NBS;
        b.x = V3<T>(s[0], s[1], s[2]);
        b.q = Q4<T>(s[3], s[4], s[5], s[6]);
        b.v = V3<T>(s[7], s[8], s[9]);
        b.w = V3<T>(s[10], s[11], s[12]);
        b.B = M3<T>::fromQuat(b.q);
    }
    const T invMass = M.body(BF_INVMASS, bidx);
    const V3<T> invI(M.body(BF_INVIX, bidx), M.body(BF_INVIY, bidx), M.body(BF_INVIZ, bidx));
    const V3<T> inertia(M.body(BF_IX, bidx), M.body(BF_IY, bidx), M.body(BF_IZ, bidx));
    const T gfy = M.body(BF_GFY, bidx);
    const T mu = M.body(BF_MU, bidx);
    const T restitutionCombined = M.body(BF_RESTITUTION, bidx) * (T)a.ground_restitution;
    const int shapeBegin = M.bodyi(BI_SHAPE_BEGIN, bidx), shapeCount = M.bodyi(BI_SHAPE_COUNT, bidx);
    const int incBegin = M.bodyi(BI_INC_BEGIN, bidx), incCount = M.bodyi(BI_INC_COUNT, bidx);
    const bool gyro = M.bodyi(BI_GYRO, bidx) != 0;
    const bool dynamic = hasBody && invMass != T(0);
    if (!(kFast && TGB_FACTORED_IINV)) b.Iinv = b.B.conjugateDiag(invI);

    auto publishPose = [&]() {
        if (hasBody) {
            /* 128-bit stores / loads: a body's record is 12 scalars = 3 (float) or 6 (double) vectors */
            T rec[12] = {b.x.x, b.x.y, b.x.z, b.B.m[0][0], b.B.m[0][1], b.B.m[0][2], b.B.m[1][0], b.B.m[1][1], b.B.m[1][2],
                         b.B.m[2][0], b.B.m[2][1], b.B.m[2][2]};
            vec_t* dst = reinterpret_cast<vec_t*>(sPose + 12 * lane);
#pragma unroll
            for (int i = 0; i < 12 / VN; ++i) dst[i] = *reinterpret_cast<vec_t*>(&rec[i * VN]);
        }
    };
    /* world inverse inertia times a vector.  Double keeps the matrix (the oracle's operations); single precision applies it in factored
       form, B (invI o (B^T v)), and does not hold the nine words of the matrix in registers across the step -- the kernel sits at the
       255-register limit and every value kept alive there is a spilled loop invariant somewhere else. */
    auto applyIinv = [&](const V3<T>& v) -> V3<T> {
        if (kFast && TGB_FACTORED_IINV) {
            const V3<T> l = b.B.tmul(v);
            return b.B * V3<T>(invI.x * l.x, invI.y * l.y, invI.z * l.z);
        }
        return b.Iinv * v;
    };
    publishPose();
    if (a.n_pairs > 0 && bodyWarp) {
        /* every pair is tested on the first step of a launch */
        for (int i = lane; i < a.n_pairs + LPE; i += LPE) sPairSlack[i] = T(0);
    }
    if (WIDE) __syncthreads(); else __syncwarp();
    /* a frame tied to a static body never moves: static bodies then stay bit-exactly in place */
    const bool frameMoves = M.body(BF_INVMASS, a.frame_body) != T(0);

    /* row column of this body: its index among the dynamic bodies of the environment when static ones have no rows */
    int rowIdx = bidx;
    if (OBST) {
        const unsigned dynMask = __ballot_sync(fullMask, dynamic);
        rowIdx = __popc(dynMask & ((1u << warpLane) - 1u) & ~((1u << groupLane0) - 1u));
    }
#define CT(slot, f) sCt[((slot) * rowBodies + rowIdx) * NCT + (f)]
    /* the impulses this body's manifold point `key` ended the last world step with, scaled by the warm starting factor (0 when new) */
    auto warmLookup = [&](T key, T& wN, T& wT) {
        wN = T(0); wT = T(0);
        if (!warmOn) return;
        const T warmFactor = (T)a.warm_factor;
        for (int j = 0; j < a.cap_contacts; ++j) {
            const T* e = sWarm + (j * rowBodies + rowIdx) * TGB_WARM_WORDS;
            if (e[0] == key) { wN = e[1] * warmFactor; wT = e[2] * warmFactor; }
        }
    };

    /* ---- CPG: CPGNode::updateDTs for this lane's nodes (CPGNode.cpp:71-88). Every lane of the warp calls this the same
       number of times (it synchronises); `act` masks groups that have nothing to do. sin(phi_j - phi_i - theta) is
       expanded over the published (sin phi, cos phi) of the nodes and the precomputed (cos theta, sin theta). */
    /* the coupling lists are read in every stage of the integrator: short ones are staged into shared memory behind the model's
       ints (TetraSpine: 205 words, +4 %); long ones stay in global memory / L1 -- for the 12-vertebra spine (several thousand words)
       the extra shared memory costs resident CTAs: 9.7e6 -> 6.5e6 env-steps/s when staged */
    int* sTopo = sModelI + ((a.model_i_words + 3) & ~3);
    const bool stageTopo = cpgOn && tgb_cpg_topo_staged(a);
    if (stageTopo) {
        const int topoWords = (C + 1) + 2 * a.cpg_n_couplings;
        for (int i = tid; i < topoWords; i += blockDim.x) sTopo[i] = a.cpg_topo[i];
        __syncthreads();
    }
    const int* topoBegin = stageTopo ? sTopo : a.cpg_topo;
    const int* topoTarget = topoBegin + (C + 1);
    const int* topoPair = topoTarget + a.cpg_n_couplings;
    /* CPG state, stage state and the seven slopes live in planes [variable][node] (consecutive lanes on consecutive doubles: the
       stride of three doubles of the node-major layout put LDS.64 / STS.64 at 50 % excessive wavefronts, ncu round 2) */
    auto cpgRhs = [&](const double* xin, double* kout, bool act) {
        if (act)
            for (int i = lane; i < C; i += LPE) {
                double sn, cn;
                sincos(xin[i], &sn, &cn);
                sTrig[i] = sn; sTrig[C + i] = cn; sTrig[2 * C + i] = xin[C + i];
            }
        __syncwarp();
        if (act) {
            const double d = a.cpg_desc;
            const bool inRange = d >= a.cpg_d_min && d <= a.cpg_d_max;
            const double freq = inRange ? sCtrl[0] * d + sCtrl[0] : 0.0;   /* nodeEquation(d, c0, c1) with c0 = c1 */
            const double radius = inRange ? sCtrl[1] * d + sCtrl[1] : 0.0;
            for (int i = lane; i < C; i += LPE) {
                const double si = sTrig[i], ci = sTrig[C + i];
                double phiDot = 2 * 3.14159265358979323846 * freq;
                for (int k = topoBegin[i]; k < topoBegin[i + 1]; ++k) {
                    const int j = topoTarget[k], p = topoPair[k];
                    const double sj = sTrig[j], cj = sTrig[C + j], rj = sTrig[2 * C + j];
                    const double sd = sj * ci - cj * si, cd = cj * ci + sj * si;
                    phiDot += sPair[p] * rj * (sd * sPair[a.cpg_n_pairs + p] - cd * sPair[2 * a.cpg_n_pairs + p]);
                }
                const double r = xin[C + i], rd = xin[2 * C + i];
                kout[i] = phiDot;
                kout[C + i] = rd;
                kout[2 * C + i] = a.cpg_r_const * (a.cpg_r_const / 4 * (radius - r) - rd);
            }
        }
        __syncwarp();
    };
    /* CPGEquations::update (CPGEquations.cpp:231-267) = Boost.Odeint integrate(): controlled Dormand-Prince 5(4), FSAL,
       eps_abs = eps_rel = 1e-6 (see oracle/tg_odeint.h for the restated algorithm). `want` = this group updates now.
       The six stages run as a loop over the tableau (tgb_dopri_a) with ONE copy of the right-hand side: unrolled, the eight copies
       of it (each with a double-precision sincos) made the kernel's loop larger than the instruction cache serves -- ncu round 2:
       35 % of the stall samples of the TetraSpine population evaluation were "no instruction".  Same operations in the same order
       (the fifth-order solution's k2 term, absent in odeint, is a product with 0 added to a finite sum: exact). */
    auto cpgUpdate = [&](double tEnd, bool want) {
        const double eps = 2.220446049250313e-16;
        double dtc = tEnd <= 0.1 ? tEnd : 0.1;
        double t = 0.0;
        int numSteps = 0;
        double* k1 = sK; double* k3 = sK + 6 * C; double* k4 = sK + 9 * C; double* k5 = sK + 12 * C;
        double* k6 = sK + 15 * C; double* k7 = sK + 18 * C;
        cpgRhs(sX, k1, want);
        numSteps += 1;
        int trials = 0;
        bool active = want && (tEnd - t > eps);
        while (__any_sync(fullMask, active)) {
            if (active && trials == 0 && (t + dtc) - tEnd > eps) dtc = tEnd - t;
            for (int stg = 0; stg < 6; ++stg) {
                if (active)
                    for (int i = lane; i < C; i += LPE) {
#pragma unroll
                        for (int var = 0; var < 3; ++var) {
                            const int v = var * C + i;
                            double acc = 1.0 * sX[v];
                            for (int j = 0; j <= stg; ++j) acc = acc + (dtc * tgb_dopri_a[stg][j]) * sK[3 * C * j + v];
                            sXt[v] = acc;
                        }
                    }
                cpgRhs(sXt, sK + 3 * C * (stg + 1), active);
            }
            const double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0, c6 = 11.0 / 84.0;
            const double dc1 = c1 - 5179.0 / 57600.0, dc3 = c3 - 7571.0 / 16695.0, dc4 = c4 - 393.0 / 640.0,
                         dc5 = c5 - (-92097.0 / 339200.0), dc6 = c6 - 187.0 / 2100.0, dc7 = -1.0 / 40.0;
            if (active) numSteps += 6;
            /* error estimate, maximum over the group's nodes */
            double err = 0.0;
            if (active) for (int i = lane; i < C; i += LPE) {
#pragma unroll
                for (int var = 0; var < 3; ++var) {
                    const int v = var * C + i;
                    const double xe = (dtc * dc1) * k1[v] + (dtc * dc3) * k3[v] + (dtc * dc4) * k4[v] + (dtc * dc5) * k5[v] + (dtc * dc6) * k6[v] + (dtc * dc7) * k7[v];
                    const double e = fabs(xe) / (1e-6 + 1e-6 * (1.0 * fabs(sX[v]) + (1.0 * dtc) * fabs(k1[v])));
                    err = fmax(err, e);
                }
            }
#pragma unroll
            for (int off = LPE / 2; off > 0; off >>= 1) err = fmax(err, __shfl_xor_sync(fullMask, err, off));
            if (active) {
                if (err > 1.0) {
                    dtc *= fmax(9.0 / 10.0 * pow(err, -1.0 / (4.0 - 1.0)), 1.0 / 5.0);
                    trials += 1;
                    if (trials >= 1000) { statusAcc |= TGB_STATUS_CPG_INEFFICIENT; active = false; } /* odeint throws overflow_error */
                } else {
                    if (err < 0.5) {
                        err = fmax(pow(5.0, -5.0), err);
                        t += dtc;
                        dtc *= 9.0 / 10.0 * pow(err, -1.0 / 5.0);
                    } else {
                        t += dtc;
                    }
                    for (int i = lane; i < C; i += LPE) {
#pragma unroll
                        for (int var = 0; var < 3; ++var) { const int v = var * C + i; sX[v] = sXt[v]; k1[v] = k7[v]; }
                    }
                    trials = 0;
                    active = tEnd - t > eps;
                }
            }
            __syncwarp();
        }
        if (want) {
            for (int i = lane; i < C; i += LPE) sNodeValue[i] = sX[C + i] * cos(sX[i]);
            if (numSteps > a.cpg_max_steps) statusAcc |= TGB_STATUS_CPG_INEFFICIENT;
        }
        __syncwarp();
    };

    auto loadShape = [&](int si) {
        ShapeConst<T> sc;
        sc.radius = M.shape(SF_RADIUS, si); sc.halfLen = M.shape(SF_HALFLEN, si); sc.margin = M.shape(SF_MARGIN, si);
        sc.spos = V3<T>(M.shape(SF_POSX, si), M.shape(SF_POSY, si), M.shape(SF_POSZ, si));
        sc.r0 = V3<T>(M.shape(SF_ROT0 + 0, si), M.shape(SF_ROT0 + 3, si), M.shape(SF_ROT0 + 6, si));
        sc.r1 = V3<T>(M.shape(SF_ROT0 + 1, si), M.shape(SF_ROT0 + 4, si), M.shape(SF_ROT0 + 7, si));
        sc.sphere = M.shape(SF_KIND, si) == T(TGB_SHAPE_SPHERE);
        sc.box = M.shape(SF_KIND, si) == T(TGB_SHAPE_BOX);
        sc.halfZ = M.shape(SF_HALFZ, si);
        return sc;
    };
    const ShapeConst<T> shape0 = loadShape(shapeCount > 0 ? shapeBegin : 0);
    const bool shape0Aligned = shape0.spos.x == T(0) && shape0.spos.y == T(0) && shape0.spos.z == T(0) &&
                               shape0.r0.x == T(1) && shape0.r0.y == T(0) && shape0.r0.z == T(0) &&
                               shape0.r1.x == T(0) && shape0.r1.y == T(1) && shape0.r1.z == T(0);

    const bool doWorld = (a.phases & 1) != 0, doModels = (a.phases & 2) != 0;
    T pairClearance = T(1e30); /* smallest surface distance of two bodies' capsules this lane has seen */
    int coupledSteps = 0, coupledRecent = 0; /* steps of this launch that hold a warp up (see below): in all, and in a row up to the last one; uniform within the group */
#ifdef TGB_STATS
    const long long statT0 = clock64();
    unsigned long long wCoupled = 0, wCoupledCyc = 0, wDetectCyc = 0, wMem = 0, wMany = 0, wCableCyc = 0;
#endif
    for (int step = 0; step < a.n_steps; ++step) {
#ifdef TGB_STATS
        int statNc = 0; bool statPen = false;
        const long long statA = clock64();
#endif
        /* =============================== world step (tgSimulation.cpp:205) =============================== */
        if (doWorld) { /* warp-uniform */
        V3<T> extLin(T(0), T(0), T(0)), extAng(T(0), T(0), T(0)), linvel(T(0), T(0), T(0));
        int nc = 0;
        bool anyPen = false;
        /* ---- contact between bodies: detection from the published poses (tgb_pairs.cuh).  Every lane of the body warp takes part; the
           lanes of a group take the environment's body pairs round robin.  nPairRows stays 0 unless a pair is within the contact band. */
        int nPairRows = 0;
#ifdef TGB_STATS
        const long long statP0 = clock64();
        int statTests = 0, statEmits = 0;
#endif
        if (a.n_pairs > 0 && bodyWarp) {
            const unsigned groupMaskP = (LPE == 32) ? 0xffffffffu : (((1u << LPE) - 1u) << groupLane0);
            const int* pairTable = sModelI + a.model_pair_off;
            for (int p0 = 0; p0 < a.n_pairs; p0 += LPE) {
                const int p = p0 + lane;
                int ia = 0, ib = 0;
                bool near = false;
                PairRecord<T> rec;
                rec.valid = false;
                if (p < a.n_pairs) {
                    const int code = pairTable[p];
                    ia = code & 31; ib = code >> 5;
                }
                /* no point of either body has moved further than its published bound since the pair's distance was taken: the pair
                   cannot be within the contact band before its slack is used up.  When a lane of the warp has to test its pair, the
                   others test theirs as well -- the instructions are issued anyway, and their next test moves further away. */
                T slack = T(1);
                if (p < a.n_pairs) slack = sPairSlack[p] - (sMotion[ia] + sMotion[ib]);
                if (__any_sync(fullMask, !(slack > T(0)))) {
                    bool compound = false;
                    if (p < a.n_pairs) {
#ifdef TGB_STATS
                        ++statTests;
#endif
                        T dist;
                        near = pairNear<T>(M, sPose, ia, ib, pairClearance, dist, rec, TGB_COOP_PAIRS ? &compound : nullptr);
                        slack = (dist - pairBreakingThreshold<T>() - T(4e-3)) * T(0.999);
                    }
                    /* bodies of several shapes: their pairs of shapes are tested by all lanes of the group together, pair after pair */
                    unsigned compoundMask = __ballot_sync(fullMask, compound) & groupMaskP;
                    while (compoundMask != 0u) { /* group-uniform */
                        const int L = __ffs(compoundMask) - 1;
                        compoundMask &= compoundMask - 1u;
                        const int ja = __shfl_sync(groupMaskP, ia, L), jb = __shfl_sync(groupMaskP, ib, L);
                        T best, exact;
                        bool nr;
                        pairNearCompoundGroup<T, ModelView<T>, LPE>(M, sPose, ja, jb, lane, groupMaskP, best, exact, nr);
                        if (warpLane == L) {
                            near = nr;
                            pairClearance = exact < pairClearance ? exact : pairClearance;
                            slack = (best - pairBreakingThreshold<T>() - T(4e-3)) * T(0.999);
                        }
                    }
                }
                if (p < a.n_pairs) sPairSlack[p] = slack;
                unsigned nearMask = __ballot_sync(fullMask, near) & groupMaskP;
                if (a.pair_cap < 0) nearMask = 0u; /* tuning aid: detection only */
                while (nearMask != 0u) { /* rare; group-uniform */
                    const int L = __ffs(nearMask) - 1;
                    nearMask &= nearMask - 1u;
                    int cnt = 0;
                    if (warpLane == L) {
                        PairWarm<T> pw;
                        pw.entries = warmOn ? sWarmPair : nullptr; pw.cap = a.pair_cap; pw.factor = (T)a.warm_factor;
                        cnt = rec.valid ? emitRecordedPair<T>(M, rec, ia, ib, pw, sPairRows, nPairRows, a.pair_cap)
                                        : emitPairRows<T>(M, sPose, ia, ib, pw, sPairRows, nPairRows, a.pair_cap);
                    }
#ifdef TGB_STATS
                    ++statEmits;
#endif
                    nPairRows += __shfl_sync(groupMaskP, cnt, L);
                }
            }
        }
#ifdef TGB_STATS
        if (a.stats) {
            /* [103] cycles of the pair detection, [58] pair tests (lane count), [59] emissions */
            const long long statP1 = clock64();
            int tests = statTests;
            for (int off = 16; off > 0; off >>= 1) tests += __shfl_xor_sync(fullMask, tests, off);
            wDetectCyc += (unsigned long long)(statP1 - statP0);
            if (warpLane == 0) {
                atomicAdd(a.stats + 103, (unsigned long long)(statP1 - statP0));
                atomicAdd(a.stats + 58, (unsigned long long)tests);
                atomicAdd(a.stats + 59, (unsigned long long)statEmits);
            }
        }
#endif
        if (dynamic) {
            /* applyGravity + solver-body external impulses */
            V3<T> totalForce(T(0), gfy, T(0)), totalTorque(T(0), T(0), T(0));
            if (simpleOn && ctrlKind == TGB_CTRL_SERIALIZED_SPINE) {
                /* the load SerializedSpineControl::onStep applied after the previous step (SerializedSpineControl.cpp:336-370):
                   btRigidBody::applyForce(force, marker relative position) adds to the total force and torque of this step */
                const double t = ctrlTime;
                double scale = 0.0;
                if (t > sCtrl[4] && t < sCtrl[5]) scale = (t - sCtrl[4]) / (sCtrl[5] - sCtrl[4]);
                else if (t >= sCtrl[5] && t < sCtrl[6]) scale = 1.0;
                else if (t >= sCtrl[6] && t < sCtrl[7]) scale = (sCtrl[7] - t) / (sCtrl[7] - sCtrl[6]);
                const int nLoaded = (int)sCtrl[11];
                for (int k = 0; k < nLoaded; ++k) {
                    const double* m = sCtrl + 12 + 5 * k;
                    if ((int)m[0] != lane) continue;
                    const V3<T> f((T)((sCtrl[8] * scale) * m[4]), (T)((sCtrl[9] * scale) * m[4]), (T)((sCtrl[10] * scale) * m[4]));
                    const V3<T> rel = b.B * V3<T>((T)m[1], (T)m[2], (T)m[3]); /* abstractMarker::getRelativePosition */
                    totalForce += f;
                    totalTorque += rel.cross(f);
                }
            }
            extLin = V3<T>(T(0), (gfy * invMass) * dt, T(0));
            if (simpleOn && ctrlKind == TGB_CTRL_SERIALIZED_SPINE) {
                extLin = (totalForce * invMass) * dt;
                extAng = ((kFast && TGB_FACTORED_IINV) ? applyIinv(totalTorque) : b.Iinv.tmul(totalTorque)) * dt;
            }
            if (gyro) {
                /* btRigidBody::computeGyroscopicForce, explicit, clamped at maxGyroscopicForce */
                if (kFast) {
                    /* same quantity evaluated in the body frame: Iinv^T (w x (I_w w)) = B (invI o (wl x (I o wl))) */
                    const V3<T> wl = b.B.tmul(b.w);
                    V3<T> gfl = wl.cross(V3<T>(inertia.x * wl.x, inertia.y * wl.y, inertia.z * wl.z));
                    const T l2 = gfl.length2();
                    if (l2 > SolverConst<T>::maxGyro() * SolverConst<T>::maxGyro()) gfl *= trsqrt(l2) * SolverConst<T>::maxGyro();
                    extAng -= (b.B * V3<T>(invI.x * gfl.x, invI.y * gfl.y, invI.z * gfl.z)) * dt;
                } else {
                    M3<T> Iw = b.B.conjugateDiag(inertia);
                    V3<T> tmp = Iw * b.w;
                    V3<T> gf = b.w.cross(tmp);
                    T l2 = gf.length2();
                    if (l2 > SolverConst<T>::maxGyro() * SolverConst<T>::maxGyro()) gf *= T(1) / tsqrt(l2) * SolverConst<T>::maxGyro();
                    extAng -= b.Iinv.tmul(gf) * dt;
                }
            }
            const V3<T> Vf((T)Vx, (T)Vy, (T)Vz);
            linvel = b.v + Vf; /* world linear velocity */

            /* ---- ground contact manifold (analytic cylinder rim points against the ground surface) ---- */
            if (GROUND && (!OBST || a.ground_kind != TGB_GROUND_EMPTY)) {
                /* signed height of the frame origin above the surface, formed in double */
                const T originHeight = (T)(a.gn[0] * Ox + a.gn[1] * Oy + a.gn[2] * Oz - a.goff);
                for (int si = shapeBegin; si < shapeBegin + shapeCount; ++si) {
                    /* the first shape's constants stay in registers across the steps (a rod is a single cylinder) */
                    const ShapeConst<T> sc = si == shapeBegin ? shape0 : loadShape(si);
                    const T radius = sc.radius, halfLen = sc.halfLen, margin = sc.margin;
                    V3<T> col0, axisDir, c;
                    if (si == shapeBegin && shape0Aligned) {
                        /* cylinder frame = body frame (a plain rod): the products below reduce to columns of the basis, bit for bit */
                        col0 = V3<T>(b.B.m[0][0], b.B.m[1][0], b.B.m[2][0]);
                        axisDir = V3<T>(b.B.m[0][1], b.B.m[1][1], b.B.m[2][1]);
                        c = b.x;
                    } else {
                        /* R = B * shapeRot: only columns 0 and 1 are needed */
                        col0 = b.B * sc.r0;
                        axisDir = b.B * sc.r1;
                        c = b.B * sc.spos + b.x;
                    }
                    const T rc = radius - margin, hc = halfLen - margin;
                    const T na = gn.dot(axisDir);
                    const V3<T> perp = -(gn - axisDir * na);
                    const T pl = kFast ? tsqrtNoBranch(perp.length2()) : perp.length();
                    /* the lowest rim point is at n.c - |n.a| hc - rc |perp| (exact arithmetic; a box: minus the lowest corner's
                       offset in the face instead of rc |perp|): a shape clearly above the contact band contributes nothing (the
                       1e-3 slack covers round-off; no result changes) */
                    T faceDrop = rc * pl;
                    V3<T> boxU, boxW;
                    if (sc.box) {
                        /* the corners of the core box: around each face of the local-Y pair, starting at the corner lowest along the
                           normal.  (boxU, boxW) are the halves of the face diagonals: corner k = 0 is +boxU, 1 is +boxW, 2 is -boxU,
                           3 is -boxW, so that the candidate loop below is shared with the cylinder's rim points */
                        const V3<T> ez = col0.cross(axisDir);
                        const T hx = rc, hz = sc.halfZ - margin;
                        const V3<T> bu = gn.dot(col0) > T(0) ? -col0 : col0, bw = gn.dot(ez) > T(0) ? -ez : ez;
                        boxU = bu * hx + bw * hz;
                        boxW = bw * hz - bu * hx;
                        faceDrop = -gn.dot(boxU);
                    }
                    if (((gn.dot(c) + originHeight) - margin) - tabs(na) * hc - faceDrop > SolverConst<T>::breakingThreshold() + T(1e-3)) continue;
                    V3<T> u;
                    if (pl > T(1e-9)) u = kFast ? perp * trcp(pl) : perp / pl;
                    else u = col0;
                    V3<T> wdir = axisDir.cross(u);
                    T rr = rc; /* candidate k sits at ce +- rr u or ce +- rr wdir */
                    if (sc.box) { u = boxU; wdir = boxW; rr = T(1); }
                    /* heights of the eight rim points in closed form (two end caps x four directions), used only to drop points
                       that are clearly out of the contact band before any vector work (same 1e-3 slack as above; a candidate
                       still has to pass the exact test).  The candidates are collected in a bit mask, then visited in the
                       manifold's order (end cap major, direction minor): lanes whose contacts sit at different rim points
                       run the row set-up below together instead of one candidate slot after the other. */
                    const T hCentre = (gn.dot(c) + originHeight) - margin;
                    const T hu = rr * gn.dot(u), hw = rr * gn.dot(wdir);
                    const T skipAbove = SolverConst<T>::breakingThreshold() + T(1e-3);
                    unsigned candidates = 0;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const T hEnd = hCentre + (e == 0 ? -hc : hc) * na;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const T hk = hEnd + ((k == 0) ? hu : (k == 1) ? hw : (k == 2) ? -hu : -hw);
                            if (!(hk > skipAbove)) candidates |= 1u << (e * 4 + k);
                        }
                    }
                    if (sc.sphere) candidates &= 1u; /* all eight candidates coincide with the centre */
                    int count = 0;
                    while (candidates != 0 && count < 4) {
                        const int slot = __ffs(candidates) - 1;
                        candidates &= candidates - 1;
                        const int e = slot >> 2, k = slot & 3;
                        {
                            const V3<T> ce = c + axisDir * (e == 0 ? -hc : hc);
                            const V3<T> dir = (k == 0) ? u : (k == 1) ? wdir : (k == 2) ? -u : -wdir;
                            const V3<T> pc = ce + dir * rr;
                            const T dist = (gn.dot(pc) + originHeight) - margin;
                            if (dist <= SolverConst<T>::breakingThreshold()) {
                                /* ---- setupContactConstraint / setupFrictionConstraint for this point ---- */
                                const V3<T> r = (pc - gn * margin) - b.x;
                                const V3<T> rxn = r.cross(gn);
                                const V3<T> angN = applyIinv(rxn);
                                const T denom = invMass + gn.dot(angN.cross(r));
                                const T invDiagN = trcp(denom);
                                const V3<T> vel = linvel + b.w.cross(r);
                                const T relVel = gn.dot(vel);
                                T restitution = restitutionCombined * -relVel;
                                if (restitution <= T(0)) restitution = T(0);
                                const V3<T> lat = vel - gn * relVel;
                                const T lat2 = lat.length2();
                                V3<T> t;
                                if (lat2 > epsilonOf<T>()) t = lat * (kFast ? trsqrt(lat2) : T(1) / tsqrt(lat2));
                                else t = planeSpace1(gn);
                                const T velDotN = gn.dot(linvel + extLin) + rxn.dot(b.w + extAng);
                                T velocityError = restitution - velDotN;
                                const T penetration = dist;
                                T erp = SolverConst<T>::erp2();
                                if (penetration > SolverConst<T>::splitThreshold()) erp = SolverConst<T>::erp();
                                T positionalError = T(0);
                                if (penetration > T(0)) velocityError -= tdiv(penetration, dt);
                                else positionalError = tdiv(-penetration * erp, dt);
                                const T penetrationImpulse = positionalError * invDiagN;
                                const T velocityImpulse = velocityError * invDiagN;
                                T rhsN, rhsPen;
                                if (penetration > SolverConst<T>::splitThreshold()) { rhsN = penetrationImpulse + velocityImpulse; rhsPen = T(0); }
                                else { rhsN = velocityImpulse; rhsPen = penetrationImpulse; anyPen = anyPen || rhsPen != T(0); }
                                const V3<T> rxt = r.cross(t);
                                const V3<T> angT = applyIinv(rxt);
                                const T denomT = invMass + t.dot(angT.cross(r));
                                const T invDiagT = trcp(denomT);
                                const T relT = t.dot(linvel + extLin) + rxt.dot(b.w);
                                const T rhsT = (T(0) - relT) * invDiagT;
                                CT(nc, CT_RXN + 0) = rxn.x; CT(nc, CT_RXN + 1) = rxn.y; CT(nc, CT_RXN + 2) = rxn.z;
                                CT(nc, CT_ANGN + 0) = angN.x; CT(nc, CT_ANGN + 1) = angN.y; CT(nc, CT_ANGN + 2) = angN.z;
                                CT(nc, CT_INVDIAGN) = invDiagN; CT(nc, CT_RHSN) = rhsN; CT(nc, CT_RHSPEN) = rhsPen;
                                const T key = (T)((si - shapeBegin) * 8 + slot);
                                T warmN, warmT;
                                warmLookup(key, warmN, warmT);
                                CT(nc, CT_LAMN) = warmN; CT(nc, CT_LAMPUSH) = T(0);
                                if (warmOn) CT(nc, CT_KEY) = key;
                                CT(nc, CT_T + 0) = t.x; CT(nc, CT_T + 1) = t.y; CT(nc, CT_T + 2) = t.z;
                                CT(nc, CT_RXT + 0) = rxt.x; CT(nc, CT_RXT + 1) = rxt.y; CT(nc, CT_RXT + 2) = rxt.z;
                                CT(nc, CT_ANGT + 0) = angT.x; CT(nc, CT_ANGT + 1) = angT.y; CT(nc, CT_ANGT + 2) = angT.z;
                                CT(nc, CT_INVDIAGT) = invDiagT; CT(nc, CT_RHST) = rhsT; CT(nc, CT_LAMT) = warmT;
                                if constexpr (OBST) {
                                    CT(nc, CT_NRM + 0) = gn.x; CT(nc, CT_NRM + 1) = gn.y; CT(nc, CT_NRM + 2) = gn.z; CT(nc, CT_MU) = mu;
                                }
                                ++nc;
                                ++count;
                            }
                        }
                    }
                }
            }

            /* ---- static obstacles: sample points of this body's shapes against the box shapes of static bodies (same scheme as
               the oracle's generateObstacleContacts: cylinder = five stations along the axis, at each the surface point facing the
               box; sphere = centre; box = eight corners; at most 4 points per pair of shapes, rows capped at cap_contacts) ---- */
            if constexpr (OBST) {
                const int* obst = sModelI + a.model_obst_off;
                const T bodyFriction = M.body(BF_FRICTION, bidx), bodyRestitution = M.body(BF_RESTITUTION, bidx);
                for (int oi = 0; oi < a.n_obst; ++oi) {
                    const int ob = obst[2 * oi], os = obst[2 * oi + 1];
                    /* the obstacle's pose as its (static) body published it, frame-relative like this body's */
                    const T* op = sPose + 12 * ob;
                    const V3<T> ox(op[0], op[1], op[2]);
                    M3<T> OB;
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) OB.m[i][j] = op[3 + i * 3 + j];
                    const ShapeConst<T> oc = loadShape(os);
                    const V3<T> e0 = OB * oc.r0, e1 = OB * oc.r1, e2 = e0.cross(e1);
                    const V3<T> co = OB * oc.spos + ox;
                    const V3<T> oh(oc.radius, oc.halfLen, oc.halfZ);
                    T obstMu = bodyFriction * M.body(BF_FRICTION, ob);
                    obstMu = obstMu < T(-10) ? T(-10) : (obstMu > T(10) ? T(10) : obstMu);
                    const T obstRestitution = bodyRestitution * M.body(BF_RESTITUTION, ob);
                    for (int si = shapeBegin; si < shapeBegin + shapeCount; ++si) {
                        const ShapeConst<T> sc = si == shapeBegin ? shape0 : loadShape(si);
                        const V3<T> col0 = b.B * sc.r0, axisDir = b.B * sc.r1, c = b.B * sc.spos + b.x;
                        const T margin = sc.margin, rc = sc.radius - margin, hc = sc.halfLen - margin;
                        const int nPoints = sc.sphere ? 1 : sc.box ? 8 : 5;
                        {
                            /* every sample point lies within `reach` of the shape's centre and the distance to the box is
                               1-Lipschitz: a pair whose centre is further than that from the contact band has no contact (1e-3 of
                               slack for round-off; no result changes) */
                            const T hz = sc.box ? sc.halfZ - margin : T(0);
                            const T reach = sc.sphere ? T(0) : tsqrt(rc * rc + hc * hc + hz * hz);
                            V3<T> unused;
                            if (boxDistance<T, kFast>(c, co, e0, e1, e2, oh, unused) - reach - margin > SolverConst<T>::breakingThreshold() + T(1e-3)) continue;
                        }
                        int count = 0;
                        for (int k = 0; k < nPoints && count < 4 && nc < a.cap_contacts; ++k) {
                            V3<T> pc;
                            if (sc.sphere) {
                                pc = c;
                            } else if (sc.box) {
                                const V3<T> ez = col0.cross(axisDir);
                                const T hx = rc, hz = sc.halfZ - margin;
                                pc = c + col0 * ((k & 1) ? hx : -hx) + axisDir * ((k & 2) ? hc : -hc) + ez * ((k & 4) ? hz : -hz);
                            } else {
                                const V3<T> pa = c + axisDir * (hc * (T(0.5) * (T)(k - 2)));
                                V3<T> g;
                                /* the surface point lies rc from the station and the distance is 1-Lipschitz: a station further
                                   than that from the contact band has no contact (1e-3 of slack; no result changes) */
                                if (boxDistance<T, kFast>(pa, co, e0, e1, e2, oh, g) - rc - margin > SolverConst<T>::breakingThreshold() + T(1e-3)) continue;
                                const V3<T> radial = g - axisDir * axisDir.dot(g);
                                if (kFast) {
                                    const T rl2 = radial.length2();
                                    pc = rl2 > T(1e-18) ? pa - radial * (rc * trsqrt(rl2)) : pa;
                                } else {
                                    const T rl = radial.length();
                                    pc = rl > T(1e-9) ? pa - radial * (rc / rl) : pa;
                                }
                            }
                            V3<T> n;
                            const T dist = boxDistance<T, kFast>(pc, co, e0, e1, e2, oh, n) - margin;
                            if (dist <= SolverConst<T>::breakingThreshold()) {
                                const V3<T> r = (pc - n * margin) - b.x;
                                const V3<T> rxn = r.cross(n);
                                const V3<T> angN = applyIinv(rxn);
                                const T denom = invMass + n.dot(angN.cross(r));
                                const T invDiagN = kFast ? trcp(denom) : T(1) / denom;
                                const V3<T> vel = linvel + b.w.cross(r);
                                const T relVel = n.dot(vel);
                                T restitution = obstRestitution * -relVel;
                                if (restitution <= T(0)) restitution = T(0);
                                const V3<T> lat = vel - n * relVel;
                                const T lat2 = lat.length2();
                                V3<T> t;
                                if (lat2 > epsilonOf<T>()) t = lat * (kFast ? trsqrt(lat2) : T(1) / tsqrt(lat2));
                                else t = planeSpace1(n);
                                const T velDotN = n.dot(linvel + extLin) + rxn.dot(b.w + extAng);
                                T velocityError = restitution - velDotN;
                                const T penetration = dist;
                                T erp = SolverConst<T>::erp2();
                                if (penetration > SolverConst<T>::splitThreshold()) erp = SolverConst<T>::erp();
                                T positionalError = T(0);
                                if (penetration > T(0)) velocityError -= tdiv(penetration, dt);
                                else positionalError = tdiv(-penetration * erp, dt);
                                const T penetrationImpulse = positionalError * invDiagN;
                                const T velocityImpulse = velocityError * invDiagN;
                                T rhsN, rhsPen;
                                if (penetration > SolverConst<T>::splitThreshold()) { rhsN = penetrationImpulse + velocityImpulse; rhsPen = T(0); }
                                else { rhsN = velocityImpulse; rhsPen = penetrationImpulse; anyPen = anyPen || rhsPen != T(0); }
                                const V3<T> rxt = r.cross(t);
                                const V3<T> angT = applyIinv(rxt);
                                const T denomT = invMass + t.dot(angT.cross(r));
                                const T invDiagT = kFast ? trcp(denomT) : T(1) / denomT;
                                const T relT = t.dot(linvel + extLin) + rxt.dot(b.w);
                                const T rhsT = (T(0) - relT) * invDiagT;
                                CT(nc, CT_RXN + 0) = rxn.x; CT(nc, CT_RXN + 1) = rxn.y; CT(nc, CT_RXN + 2) = rxn.z;
                                CT(nc, CT_ANGN + 0) = angN.x; CT(nc, CT_ANGN + 1) = angN.y; CT(nc, CT_ANGN + 2) = angN.z;
                                CT(nc, CT_INVDIAGN) = invDiagN; CT(nc, CT_RHSN) = rhsN; CT(nc, CT_RHSPEN) = rhsPen;
                                const T key = (T)(TGB_WARM_OBST + (oi * 64 + (si - shapeBegin)) * 8 + k);
                                T warmN, warmT;
                                warmLookup(key, warmN, warmT);
                                CT(nc, CT_LAMN) = warmN; CT(nc, CT_LAMPUSH) = T(0);
                                if (warmOn) CT(nc, CT_KEY) = key;
                                CT(nc, CT_T + 0) = t.x; CT(nc, CT_T + 1) = t.y; CT(nc, CT_T + 2) = t.z;
                                CT(nc, CT_RXT + 0) = rxt.x; CT(nc, CT_RXT + 1) = rxt.y; CT(nc, CT_RXT + 2) = rxt.z;
                                CT(nc, CT_ANGT + 0) = angT.x; CT(nc, CT_ANGT + 1) = angT.y; CT(nc, CT_ANGT + 2) = angT.z;
                                CT(nc, CT_INVDIAGT) = invDiagT; CT(nc, CT_RHST) = rhsT; CT(nc, CT_LAMT) = warmT;
                                CT(nc, CT_NRM + 0) = n.x; CT(nc, CT_NRM + 1) = n.y; CT(nc, CT_NRM + 2) = n.z; CT(nc, CT_MU) = obstMu;
                                ++nc;
                                ++count;
                            }
                        }
                    }
                }
            }
#ifdef TGB_STATS
            statNc = nc; statPen = anyPen;
#endif
        } /* dynamic: external impulses and the rows against the static world */
            /* a step that holds its warp up: rows between two bodies, or a body with three or more contact points (a rod standing on
               an end cap) -- environments in which such steps last are given a warp of their own by the host (tgb_capi.cu, tail split) */
            if (a.coupled_steps != nullptr) {
                const unsigned groupMaskH = (LPE == 32) ? 0xffffffffu : (((1u << LPE) - 1u) << groupLane0);
                const unsigned manyMask = __ballot_sync(fullMask, nc >= 3); /* (by every lane: not behind the || of a lane's own flag) */
                const int heavy = (nPairRows > 0 || (manyMask & groupMaskH) != 0u) ? 1 : 0;
                coupledSteps += heavy;
                coupledRecent = heavy ? coupledRecent + 1 : 0; /* heavy steps in a row up to now: what the environment is like at the end of the launch says most about the next one */
            }
            /* ---- projected Gauss-Seidel: split-impulse rows, then velocity rows ---- */
            V3<T> dLin(T(0), T(0), T(0)), dAng(T(0), T(0), T(0)), push(T(0), T(0), T(0)), turn(T(0), T(0), T(0));
            bool solved = false;
            if (a.n_pairs > 0 && bodyWarp && __any_sync(fullMask, nPairRows > 0)) {
                /* warp-uniform: environments of this warp have rows between two bodies.  Their lanes solve together (tgb_pairs.cuh);
                   the other environments keep their per-lane solvers below. */
#ifdef TGB_STATS
                const long long statS0 = clock64();
#endif
                CoupledBody<T> cb;
                cb.pos = b.x; cb.gnv = gn; cb.mine = dynamic; cb.invMassV = dynamic ? invMass : T(0); cb.muV = mu;
                cb.inertia.factored = kFast && TGB_FACTORED_IINV;
                cb.inertia.B = (kFast && TGB_FACTORED_IINV) ? b.B : b.Iinv;
                cb.inertia.invI = dynamic ? invI : V3<T>(T(0), T(0), T(0));
                if (!(kFast && TGB_FACTORED_IINV) && !dynamic) {
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) cb.inertia.B.m[i][j] = T(0);
                }
                const CoupledOut<T> co = solveCoupledWarp<T, kFast, OBST, LPE>(sCt + rowIdx * NCT, rowBodies * NCT, nc, anyPen, sPairRows,
                                                                               warmOn ? sWarmPair : nullptr, nPairRows, lane, groupLane0, cb,
                                                                               kFast ? b.v : linvel, b.w, extLin, extAng, dt, a.iterations
#ifdef TGB_STATS
                                                                               , a.stats
#endif
                                                                               );
                solved = co.solved;
                if (solved) { dLin = co.o.dLin; dAng = co.o.dAng; push = co.o.push; turn = co.o.turn; }
#ifdef TGB_STATS
                __syncwarp();
                const unsigned long long statMem = co.statMem ? 1ULL : 0ULL;
                const unsigned statCoupledLanes = __ballot_sync(fullMask, nPairRows > 0); /* by every lane: lane 0 alone must not call it */
                wCoupled += 1; wCoupledCyc += (unsigned long long)(clock64() - statS0); wMem += statMem;
                if (a.stats && warpLane == 0) {
                    /* [55] warp-steps with a coupled environment, [56] of which with one in the shared-memory solver, [57] coupled
                       environments summed over warp-steps, [104] cycles of the coupled solve */
                    atomicAdd(a.stats + 55, 1ULL);
                    if (statMem) atomicAdd(a.stats + 56, 1ULL);
                    atomicAdd(a.stats + 57, (unsigned long long)(__popc(statCoupledLanes) / LPE));
                    atomicAdd(a.stats + 104, (unsigned long long)(clock64() - statS0));
                    atomicAdd(a.stats + 110, (unsigned long long)(statS0 - statA)); /* body phase before the coupled solve, coupled warp-steps only */
                }
#endif
            }
            if (solved) {
                /* done above */
            } else if (!dynamic) {
                /* nothing to solve */
            } else if (kFast && TGB_SINGLE_CONTACT_PGS && nc == 1) {
                /* in worlds with obstacles the single row carries its own normal and friction coefficient */
                const V3<T> gn1 = OBST ? V3<T>(CT(0, CT_NRM + 0), CT(0, CT_NRM + 1), CT(0, CT_NRM + 2)) : gn;
                const T mu1 = OBST ? CT(0, CT_MU) : mu;
                /* Single precision, one contact point (by far the commonest case: a rod end touching the ground): the matrix form
                   of the solver below for the two rows (normal, friction) of that point, straight-line and branch-free.  The
                   relative velocity change of each row is carried along: dvN = ANN lamN + ANT lamT, dvT = ANT lamN + ATT lamT. */
                const V3<T> rxn(CT(0, CT_RXN + 0), CT(0, CT_RXN + 1), CT(0, CT_RXN + 2));
                const V3<T> angN(CT(0, CT_ANGN + 0), CT(0, CT_ANGN + 1), CT(0, CT_ANGN + 2));
                const V3<T> t(CT(0, CT_T + 0), CT(0, CT_T + 1), CT(0, CT_T + 2));
                const V3<T> rxt(CT(0, CT_RXT + 0), CT(0, CT_RXT + 1), CT(0, CT_RXT + 2));
                const V3<T> angT(CT(0, CT_ANGT + 0), CT(0, CT_ANGT + 1), CT(0, CT_ANGT + 2));
                const T invDiagN = CT(0, CT_INVDIAGN), rhsN = CT(0, CT_RHSN), rhsPen = CT(0, CT_RHSPEN);
                const T invDiagT = CT(0, CT_INVDIAGT), rhsT = CT(0, CT_RHST);
                const V3<T> bn = gn1 * invMass, bt = t * invMass;
                /* With one point the two rows couple through ANT = J_N M^-1 J_T^T only, and 1 / invDiag is the row's own diagonal
                   entry, so an iteration  lam += rhs - (A lam)_row * invDiag  collapses to
                       lamN <- max(0, rhsN - kN lamT),   lamT <- clamp(rhsT - kT lamN, -mu lamN, mu lamN)
                   with kN = ANT invDiagN, kT = ANT invDiagT: the same fixed-point iteration in exact arithmetic with a dependent
                   chain of five operations per iteration instead of eleven.  The split-impulse row is alone in its system:
                   every iteration leaves it at max(0, rhsPen). */
                const T ANT = gn1.dot(bt) + rxn.dot(angT);
                const T kN = ANT * invDiagN, kT = ANT * invDiagT;
                if (rhsPen != T(0) && a.iterations > 0) {
                    const T lamPush = rhsPen > T(0) ? rhsPen : T(0);
                    push = bn * lamPush;
                    turn = angN * lamPush;
                }
                /* warm starting: the normal row's own diagonal cancels its initial impulse out of the iteration above; the friction
                   impulse enters through the first normal update (and stays if the normal impulse never becomes positive) */
                T lamN = T(0), lamT = T(0);
                if (warmOn) { lamN = CT(0, CT_LAMN); lamT = CT(0, CT_LAMT); }
                for (int it = 0; it < a.iterations; ++it) {
                    const T sumN = rhsN - kN * lamT;
                    lamN = sumN < T(0) ? T(0) : sumN;
                    /* friction row: skipped (impulse kept) while the normal impulse is zero */
                    const T hi = mu1 * lamN;
                    T sumT = rhsT - kT * lamN;
                    sumT = sumT < -hi ? -hi : (sumT > hi ? hi : sumT);
                    lamT = lamN > T(0) ? sumT : lamT;
                }
                dLin = bn * lamN + bt * lamT;
                dAng = angN * lamN + angT * lamT;
                if (warmOn) { CT(0, CT_LAMN) = lamN; CT(0, CT_LAMT) = lamT; }
            } else if (OBST) {
                if (kFast && TGB_MATRIX_PGS && TGB_MATRIX4 && nc > 0 && nc <= 4) {
                    const ContactSolveOut<T> o = solveContactsMatrix4<T, true>(sCt + rowIdx * NCT, rowBodies * NCT, nc, a.iterations, anyPen, gn, invMass, mu);
                    dLin = o.dLin; dAng = o.dAng; push = o.push; turn = o.turn;
                } else if (nc > 0) pgsRowsWithNormals<T>(sCt + rowIdx * NCT, rowBodies * NCT, nc, a.iterations, anyPen, invMass, dLin, dAng, push, turn);
            } else if (nc > 0) {
                const ContactSolveOut<T> o = solveContactsGeneral<T, kFast>(sCt + rowIdx * NCT, rowBodies * NCT, nc, a.iterations, anyPen, gn, invMass, mu);
                dLin = o.dLin; dAng = o.dAng; push = o.push; turn = o.turn;
            }

            /* ---- the persistent manifold of the next step: this step's rows with their final impulses (every solver above leaves them
               in the rows); slots beyond them are emptied ---- */
            if (warmOn) {
                if (dynamic) {
                    for (int j = 0; j < a.cap_contacts; ++j) {
                        T* e = sWarm + (j * rowBodies + rowIdx) * TGB_WARM_WORDS;
                        if (j < nc) { e[0] = CT(j, CT_KEY); e[1] = CT(j, CT_LAMN); e[2] = CT(j, CT_LAMT); }
                        else e[0] = T(-1);
                    }
                }
                if (a.n_pairs > 0 && bodyWarp) /* nPairRows: uniform within the group */
                    for (int p = nPairRows + lane; p < warmPairCap; p += LPE) sWarmPair[p * TGB_WARM_WORDS] = T(-1);
            }
            /* ---- write back, split-impulse correction, integrateTransforms ---- */
            if (dynamic) {
                b.v = (b.v + dLin) + extLin;
                b.w = (b.w + dAng) + extAng;
                if (push.nonzero() || turn.nonzero()) integrateTransform(b.x, b.q, push, turn * SolverConst<T>::turnErp(), dt);
                integrateTransform(b.x, b.q, b.v, b.w, dt);
                b.B = M3<T>::fromQuat(b.q);
                if (!(kFast && TGB_FACTORED_IINV)) b.Iinv = b.B.conjugateDiag(invI);
                if (a.n_pairs > 0) {
                    /* how far a point of this body's shapes can have moved in this step, relative to the env frame: translation by
                       (v + push) dt, rotation by at most (|w| + turnErp |turn|) dt about the body origin (1-norms of the rare split-impulse terms: upper bounds) */
                    const T lin1 = tsqrtNoBranch(b.v.length2()) + (tabs(push.x) + tabs(push.y) + tabs(push.z));
                    const T ang1 = tsqrtNoBranch(b.w.length2()) + SolverConst<T>::turnErp() * (tabs(turn.x) + tabs(turn.y) + tabs(turn.z));
                    sMotion[lane] = (lin1 + M.body(BF_BOUND, bidx) * ang1) * (dt * T(1.001));
                }
            }
        } /* doWorld */

        /* ---- env frame: advance the origin with the frame velocity, then re-centre on the frame body ---- */
        if (frameMoves && doWorld && bodyWarp) {
            Ox += Vx * a.dt; Oy += Vy * a.dt; Oz += Vz * a.dt;
            const int src = groupLane0 + a.frame_body;
            const T sx = __shfl_sync(fullMask, b.x.x, src), sy = __shfl_sync(fullMask, b.x.y, src), sz = __shfl_sync(fullMask, b.x.z, src);
            const T svx = __shfl_sync(fullMask, b.v.x, src), svy = __shfl_sync(fullMask, b.v.y, src), svz = __shfl_sync(fullMask, b.v.z, src);
            Ox += (double)sx; Oy += (double)sy; Oz += (double)sz;
            Vx += (double)svx; Vy += (double)svy; Vz += (double)svz;
            b.x -= V3<T>(sx, sy, sz);
            b.v -= V3<T>(svx, svy, svz);
        }
        publishPose();
        if (WIDE) __syncthreads(); else __syncwarp();
#ifdef TGB_STATS
        if (a.stats) {
            /* per warp and step: [0..8] histogram of the largest contact count of a lane, [16..48] histogram of the number of
               lanes with contacts, [49] warp-steps with a split-impulse row, [50] warp-steps */
            int mx = statNc;
            for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(fullMask, mx, off));
            const int touching = __popc(__ballot_sync(fullMask, statNc > 0));
            const bool pen = __any_sync(fullMask, statPen);
            wMany += mx >= 3 ? 1 : 0;
            if (warpLane == 0) {
                atomicAdd(a.stats + min(mx, 8), 1ULL);
                atomicAdd(a.stats + 16 + touching, 1ULL);
                if (pen) atomicAdd(a.stats + 49, 1ULL);
                atomicAdd(a.stats + 50, 1ULL);
            }
        }
        const long long statB = clock64();
#endif

        /* =============================== models step (tgSimulation.cpp:208-211) =============================== */
        if (!doModels) continue; /* warp-uniform: a world-half-only launch (host observers run between the halves) */
        if (sineOn) ctrlTime += a.dt; /* EscapeController::m_totalTime */
        if (simpleOn && ctrlKind == TGB_CTRL_SERIALIZED_SPINE) {
            /* SerializedSpineControl::onStep, SerializedSpineControl.cpp:264-278: simTime advances in jumps of 1 / updateFrequency */
            frameSpare += a.dt;
            if (frameSpare >= 1.0 / sCtrl[3]) { ctrlTime += frameSpare; frameSpare = 0.0; }
        }
        bool cpgCableControl = false;
        double cpgCableDt = 0.0;
        if (cpgOn) {
            /* BaseSpineCPGControl::onStep, BaseSpineCPGControl.cpp:253-279 */
            cpgUpdateTime += a.dt;
            const bool want = cpgUpdateTime >= a.cpg_control_time;
            if (__any_sync(fullMask, want)) cpgUpdate(cpgUpdateTime, want);
            if (want) cpgUpdateTime = 0.0;
            if (hasBody && lane == a.cpg_score_body) {
                const double h = Oy + (double)b.x.y;
                if (h > a.cpg_height_hi || h < a.cpg_height_lo) statusAcc |= TGB_STATUS_HEIGHT; /* bogus */
            }
            /* tgCPGActuatorControl::onStep timing, tgCPGActuatorControl.cpp:79-100 (identical for every cable) */
            cpgCableTime += a.dt;
            cpgCableControl = cpgCableTime >= a.cpg_cable_step;
            cpgCableDt = cpgCableTime;
            if (cpgCableControl) cpgCableTime = 0.0;
        }
        /* Cables of this lane are processed CU at a time in three passes (loads, arithmetic, stores) over independent
           register sets, so that the loads and the dependent chains of different cables overlap instead of queueing. */
        {
        /* the cable role's view of its environment (identical to the body role's in the narrow layout) */
        T* const sState = cStateP; T* const sPose = cPoseP; T* const sJ = cJP;
        double* const sCtrl = cCtrlP; double* const sCum = cCumP; double* const sSine = cSineP; double* const sRot = cRotP;
        const int env = cEnv; const bool live = cLive;
        for (int c0 = cLane; c0 < C; c0 += CU * CL) {
            CableRegs<T> cr[CU];
#pragma unroll
            for (int u = 0; u < CU; ++u) {
                CableRegs<T>& q = cr[u];
                const int cu = c0 + u * CL;
                q.valid = cu < C;
                const int c = q.valid ? cu : c0; /* an idle slot shadows cable c0 and never stores */
                q.c = c;
#if TGB_VEC_CREC
                T rc_[16];
                {
                    const vec_t* src = reinterpret_cast<const vec_t*>(M.crec + c * TGB_CREC);
#pragma unroll
                    for (int i = 0; i < 16 / VN; ++i) *reinterpret_cast<vec_t*>(&rc_[i * VN]) = src[i];
                }
                const int4 ri = *reinterpret_cast<const int4*>(M.cirec + c * TGB_CIREC);
                const int ia = ri.x, ib = ri.y;
                q.actuator = ri.z & 0xff;
                q.slots = ri.w; q.ia = ia; q.ib = ib;
                const V3<T> la(rc_[0], rc_[1], rc_[2]);
                const V3<T> lb(rc_[4], rc_[5], rc_[6]);
                q.k = rc_[3]; q.damp = rc_[7];
                q.tVel = rc_[8]; q.maxTens = rc_[9]; q.maxStretch = rc_[10];
                q.minAL = rc_[11]; q.minRL = rc_[12];
#else
                const int ia = M.cablei(CI_A, c), ib = M.cablei(CI_B, c);
                q.actuator = M.cablei(CI_ACTUATOR, c);
                q.slots = M.cirec[c * TGB_CIREC + 3]; q.ia = ia; q.ib = ib;
                const V3<T> la(M.cable(CF_LAX, c), M.cable(CF_LAY, c), M.cable(CF_LAZ, c));
                const V3<T> lb(M.cable(CF_LBX, c), M.cable(CF_LBY, c), M.cable(CF_LBZ, c));
                q.k = M.cable(CF_K, c); q.damp = M.cable(CF_C, c);
                q.tVel = M.cable(CF_TVEL, c); q.maxTens = M.cable(CF_MAXTENS, c); q.maxStretch = M.cable(CF_MAXSTRETCH, c);
                q.minAL = M.cable(CF_MINAL, c); q.minRL = M.cable(CF_MINRL, c);
#endif
                T pa[12], pb[12];
                {
                    const vec_t* srcA = reinterpret_cast<const vec_t*>(sPose + 12 * ia);
                    const vec_t* srcB = reinterpret_cast<const vec_t*>(sPose + 12 * ib);
#pragma unroll
                    for (int i = 0; i < 12 / VN; ++i) {
                        *reinterpret_cast<vec_t*>(&pa[i * VN]) = srcA[i];
                        *reinterpret_cast<vec_t*>(&pb[i * VN]) = srcB[i];
                    }
                }
                const V3<T> xa(pa[0], pa[1], pa[2]), xb(pb[0], pb[1], pb[2]);
                M3<T> Ba, Bb;
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) { Ba.m[i][j] = pa[3 + i * 3 + j]; Bb.m[i][j] = pb[3 + i * 3 + j]; }
                /* anchors: tgBulletSpringCableAnchor.cpp:83-87 */
                q.ra = Ba * la; q.rb = Bb * lb;
                const V3<T> p1 = q.ra + xa, p2 = q.rb + xb;
                q.dist = p2 - p1;
                q.L = kFast ? tsqrtNoBranch(q.dist.length2()) : q.dist.length();
#if TGB_VEC_CS
                T cs[TGB_NCS];
                {
                    const vec_t* src = reinterpret_cast<const vec_t*>(sState + TGB_CABLE_BASE(R) + c * TGB_NCS);
#pragma unroll
                    for (int i = 0; i < TGB_NCS / VN; ++i) *reinterpret_cast<vec_t*>(&cs[i * VN]) = src[i];
                }
#else
                const T* cs = sState + TGB_CABLE_BASE(R) + c * TGB_NCS;
#endif
                q.rest = cs[CS_REST]; q.prev = cs[CS_PREV]; q.pref = cs[CS_PREF]; q.motorVel = cs[CS_MOTORVEL];
                q.desiredTorque = T(0);
            }

            /* ---- model-level controllers (run before the children step) ---- */
            if (simpleOn && ctrlKind == TGB_CTRL_SETPOINT) {
#pragma unroll
                for (int u = 0; u < CU; ++u) {
                    CableRegs<T>& q = cr[u];
                    if (q.actuator == TGB_ACT_BASIC) {
                        const T input = (T)sCtrl[q.c];
                        if (!(input < T(0))) { q.pref = input; q.moveMotors(dt); }
                    }
                }
            } else if (simpleOn && ctrlKind == TGB_CTRL_TORQUE) {
#pragma unroll
                for (int u = 0; u < CU; ++u)
                    if (cr[u].actuator == TGB_ACT_KINEMATIC) cr[u].desiredTorque = (T)sCtrl[cr[u].c];
            } else if (simpleOn && ctrlKind == TGB_CTRL_SERIALIZED_SPINE) {
#pragma unroll
                for (int u = 0; u < CU; ++u) {
                    CableRegs<T>& q = cr[u];
                    if (q.actuator == TGB_ACT_BASIC) {
                        /* SerializedSpineControl::applyImpedanceControlGeneric, SerializedSpineControl.cpp:174-205 */
                        const double* g = sCtrl + TGB_SPINE_HEADER + 5 * q.c;
                        const T cycle = sinPhase(ctrlTime * sCtrl[2] + g[4], T(0));
                        const T target = (T)sCtrl[0] + cycle * (T)sCtrl[1];
                        /* tgImpedanceController::controlTension(tgBasicActuator&, ...), tgImpedanceController.cpp:163-185 */
                        const T vel = q.motorVel; /* velocity slot of a basic actuator = the cable's m_velocity */
                        T setTension = (T)g[0] + (T)g[1] * (q.L - (T)g[3]) + (T)g[2] * (vel - target);
                        setTension = setTension > T(0) ? setTension : T(0);
                        /* static tgTensionController::control(sca, dt, setPoint), tgTensionController.cpp:90-114 */
                        T tension = (q.L - q.rest) * q.k;
                        tension = tension < T(0) ? T(0) : tension;
                        const T delta = setTension - tension;
                        const T diff = delta / q.k;
                        T newLength = q.rest - diff;
                        newLength = newLength < T(0.1) ? T(0.1) : newLength;
                        q.pref = newLength;
                        q.moveMotors(dt);
                    }
                }
            } else if (simpleOn && ctrlKind == TGB_CTRL_TENSION) {
                const T setPoint = (T)sCtrl[0];
#pragma unroll
                for (int u = 0; u < CU; ++u) {
                    CableRegs<T>& q = cr[u];
                    if (q.actuator == TGB_ACT_BASIC) {
                        /* tgTensionController::control(dt, setPoint): tgTensionController.cpp:50-88 */
                        T tension = (q.L - q.rest) * q.k;
                        tension = tension < T(0) ? T(0) : tension;
                        const T delta = setPoint - tension;
                        const T diff = delta / q.k;
                        T newLength = q.rest - diff;
                        newLength = newLength < T(0) ? T(0) : newLength;
                        q.pref = newLength;
                        q.moveMotors(dt);
                    }
                }
            } else if (sineOn) {
                const T L0 = (T)sCtrl[32];
                const T minLength = L0 * (T(1) - T(0.50)), maxLength = L0 * (T(1) + T(0.50));
#pragma unroll
                for (int u = 0; u < CU; ++u) {
                    CableRegs<T>& q = cr[u];
                    const int c = q.c;
                    if (c < 24 && q.valid) { /* an idle unroll slot shadows cable c0: it must not advance c0's phase a second time */
                        /* EscapeController::setPreferredMuscleLengths, EscapeController.cpp:245-291 */
                        const int cluster = c < 16 ? (c >> 2) : 4 + ((c - 16) >> 1);
                        T newLength;
                        if (sineRecurrence) {
                            double* sc = sSine + c;
                            const double s0 = sc[0], c0r = sc[24], cd = sRot[2 * cluster], sd = sRot[2 * cluster + 1];
                            const double s1 = s0 * cd + c0r * sd, c1 = c0r * cd - s0 * sd;
                            sc[0] = s1; sc[24] = c1;
                            newLength = TGB_SINE_REC == 4 ? (T)(sc[48] * s1 + sc[72]) : (T)(sCtrl[cluster * 4 + 0] * s1 + sCtrl[cluster * 4 + 3]);
                        } else {
                            const bool firstOfCluster = c < 16 ? ((c & 3) == 0) : (((c - 16) & 1) == 0);
                            const double phase = (firstOfCluster && cluster > 0) ? sCum[cluster - 1] : sCum[cluster];
                            const double arg = sCtrl[cluster * 4 + 1] * ctrlTime + phase;
                            newLength = (T)sCtrl[cluster * 4 + 0] * sinPhase(arg, T(0)) + (T)sCtrl[cluster * 4 + 3];
                        }
                        if (newLength <= minLength) newLength = minLength;
                        else if (newLength >= maxLength) newLength = maxLength;
                        const bool accept = !(newLength < T(0));
                        q.pref = accept ? newLength : q.pref;
                        if (kFast && TGB_FUSE_DOUBLE_MOVE) {
                            q.moveMotors(dt, true, accept ? T(2) : T(1)); /* both moves of this step in one */
                            continue;
                        }
                        q.moveMotors(dt, accept);
                    }
                    q.moveMotors(dt); /* EscapeController.cpp:108-116: every muscle's motors move a second time */
                }
            } else if (cpgOn) {
#pragma unroll
                for (int u = 0; u < CU; ++u) {
                    CableRegs<T>& q = cr[u];
                    const double* g = a.cpg_imp + 4 * q.c;
                    if (q.actuator == TGB_ACT_BASIC) {
                        if (cpgCableControl) {
                            /* tgImpedanceController::controlTension(tgBasicActuator&, ...), tgImpedanceController.cpp:163-185 */
                            const T vel = q.motorVel; /* velocity slot of a basic actuator = the cable's m_velocity */
                            T setTension = (T)g[0] + (T)g[1] * (q.L - (T)g[3]) + (T)g[2] * (vel - (T)sNodeValue[q.c]);
                            setTension = setTension > T(0) ? setTension : T(0);
                            /* static tgTensionController::control(sca, dt, setPoint), tgTensionController.cpp:90-114 */
                            T tension = (q.L - q.rest) * q.k;
                            tension = tension < T(0) ? T(0) : tension;
                            const T delta = setTension - tension;
                            const T diff = delta / q.k;
                            T newLength = q.rest - diff;
                            newLength = newLength < T(0.1) ? T(0.1) : newLength;
                            q.pref = newLength;
                            /* tgCPGActuatorControl passes the accumulated control time, tgCPGCableControl the step's dt */
                            q.moveMotors(a.cpg_cable_control == 1 ? dt : (T)cpgCableDt);
                        } else {
                            q.moveMotors(dt);
                        }
                    } else if (q.actuator == TGB_ACT_KINEMATIC && a.cpg_cable_control == 1 && q.valid) {
                        /* tgCPGCableControl::onStep with a tgPIDController, tgCPGCableControl.cpp:63-98 */
                        double* pid = sPid + 3 * q.c;
                        T tension = (q.L - q.rest) * q.k;
                        tension = tension < T(0) ? T(0) : tension;
                        T setPoint = (T)pid[2];
                        if (cpgCableControl) {
                            /* tgImpedanceController::controlTension(tgBasicController&, ...), tgImpedanceController.cpp:118-150;
                               the actuator's velocity is motor velocity x radius (tgKinematicActuator.cpp:181-184) */
                            const T vel = q.motorVel * M.cable(CF_MRADIUS, q.c);
                            setPoint = (T)g[0] + (T)g[1] * (q.L - (T)g[3]) + (T)g[2] * (vel - (T)sNodeValue[q.c]);
                            setPoint = setPoint > T(0) ? setPoint : T(0);
                            pid[2] = (double)setPoint;
                        }
                        /* tgPIDController::control(dt, setPoint, sensorData), tgPIDController.cpp:80-109 */
                        const T prevError = (T)pid[0];
                        const T error = setPoint - tension;
                        const T intError = (T)pid[1] + (error + prevError) / T(2) * dt;
                        const T dError = (error - prevError) / dt;
                        q.desiredTorque = (T)a.cpg_pid[0] * error + (T)a.cpg_pid[1] * intError + (T)a.cpg_pid[2] * dError;
                        pid[0] = (double)error; pid[1] = (double)intError;
                    }
                }
            }

            /* ---- tgKinematicActuator::integrateRestLength, tgKinematicActuator.cpp:186-225 ---- */
#pragma unroll
            for (int u = 0; u < CU; ++u) {
                CableRegs<T>& q = cr[u];
                if (q.actuator == TGB_ACT_KINEMATIC) {
                    const int c = q.c;
                    T tension = (q.L - q.rest) * q.k;
                    tension = tension < T(0) ? T(0) : tension;
                    const T mRadius = M.cable(CF_MRADIUS, c);
                    T maxTorque = q.maxTens * mRadius * (T(1) - mRadius * tabs(q.motorVel) / q.tVel);
                    maxTorque = maxTorque < T(0) ? T(0) : maxTorque;
                    const T applied = tabs(q.desiredTorque) < maxTorque ? q.desiredTorque : q.desiredTorque / tabs(q.desiredTorque) * maxTorque;
                    const T acc = (applied - M.cable(CF_MFRICTION, c) * q.motorVel + tension * mRadius) / M.cable(CF_MINERTIA, c);
                    if (M.cablei(CI_BACKDRIVABLE, c) == 0 && acc * applied <= T(0))
                        q.motorVel = q.motorVel + acc * dt > T(0) ? T(0) : q.motorVel + acc * dt;
                    else
                        q.motorVel += acc * dt;
                    q.rest += mRadius * q.motorVel * dt;
                    q.rest = (q.rest > q.minRL) ? q.rest : q.minRL;
                }
            }

            /* ---- tgBulletSpringCable::calculateAndApplyForce, tgBulletSpringCable.cpp:90-141 ---- */
#pragma unroll
            for (int u = 0; u < CU; ++u) {
                CableRegs<T>& q = cr[u];
                const V3<T> unit = q.dist * trcp(q.L);
                const T stretch = q.L - q.rest;
                T magnitude = q.k * stretch;
                const T deltaStretch = q.L - q.prev;
                q.velocity = kFast ? deltaStretch * invDt : deltaStretch / dt;
                T damping = q.damp * q.velocity;
                if (tabs(magnitude) * T(1) < tabs(damping)) damping = (damping > T(0) ? magnitude * T(1) : -magnitude * T(1));
                magnitude += damping;
                V3<T> force(T(0), T(0), T(0));
                if (q.L > q.rest) force = unit * magnitude;
                q.prev = q.L;
                q.J = force * dt;
            }
#pragma unroll
            for (int u = 0; u < CU; ++u) {
                const CableRegs<T>& q = cr[u];
                if (!q.valid) continue;
                const int c = q.c;
                const V3<T> J = q.J;
                if (kFast) {
                    /* also publish the torque of the impulse about both bodies' centres: r x J and r x (-J) */
                    const V3<T> ta = q.ra.cross(J), tb = J.cross(q.rb);
                    /* slot of (position k in the body's list, body b): plane J at [k][(b + k) mod LPE], plane torque behind it; the
                       skew by k spreads the ends of one body (distinct k) over the banks for the scattered stores */
                    const int ka = q.slots & 0xffff, kb = q.slots >> 16;
                    float4* oa = reinterpret_cast<float4*>(sJ) + ka * LPE + ((q.ia + ka) & (LPE - 1));
                    float4* ob = reinterpret_cast<float4*>(sJ) + kb * LPE + ((q.ib + kb) & (LPE - 1));
                    const int plane = a.max_inc * LPE;
                    oa[0] = make_float4(J.x, J.y, J.z, 0.0f); oa[plane] = make_float4(ta.x, ta.y, ta.z, 0.0f);
                    ob[0] = make_float4(-J.x, -J.y, -J.z, 0.0f); ob[plane] = make_float4(tb.x, tb.y, tb.z, 0.0f);
                } else {
                    sJ[c * 3 + 0] = J.x; sJ[c * 3 + 1] = J.y; sJ[c * 3 + 2] = J.z;
                }
                if (a.max_tension != nullptr && live) {
                    /* logHistory with hist = true: the tension after this cable's step; only the running maximum is kept */
                    T tensionNow = (q.L - q.rest) * q.k;
                    tensionNow = tensionNow < T(0) ? T(0) : tensionNow;
                    T* mx = reinterpret_cast<T*>(a.max_tension) + (size_t)env * C + c;
                    if (tensionNow > *mx) *mx = tensionNow;
                }
#if TGB_VEC_CS
                T cs[TGB_NCS];
                cs[CS_REST] = q.rest; cs[CS_PREV] = q.prev; cs[CS_PREF] = q.pref;
                cs[CS_MOTORVEL] = q.actuator == TGB_ACT_KINEMATIC ? q.motorVel : q.velocity;
                vec_t* dst = reinterpret_cast<vec_t*>(sState + TGB_CABLE_BASE(R) + c * TGB_NCS);
#pragma unroll
                for (int i = 0; i < TGB_NCS / VN; ++i) dst[i] = *reinterpret_cast<vec_t*>(&cs[i * VN]);
#else
                T* cs = sState + TGB_CABLE_BASE(R) + c * TGB_NCS;
                cs[CS_REST] = q.rest; cs[CS_PREV] = q.prev; cs[CS_PREF] = q.pref;
                cs[CS_MOTORVEL] = q.actuator == TGB_ACT_KINEMATIC ? q.motorVel : q.velocity;
#endif
            }
        }
        (void)sCtrl; (void)sCum; (void)sSine; (void)sRot; (void)env; (void)live; (void)sPose; (void)sState; (void)sJ;
        }
        if (WIDE) __syncthreads(); else __syncwarp();
#ifdef TGB_STATS
        const long long statC = clock64();
#endif

        /* ---- applyImpulse of every incident cable end, in cable order (btRigidBody::applyImpulse) ---- */
        if (dynamic) {
            if (kFast) {
                /* sum the impulses and torques of the incident cable ends first, then one multiplication each */
                V3<T> sumJ(T(0), T(0), T(0)), sumT(T(0), T(0), T(0));
                const float4* o = reinterpret_cast<const float4*>(sJ);
                const int plane = a.max_inc * LPE;
                for (int k = 0; k < incCount; ++k) {
                    const int slot = k * LPE + ((lane + k) & (LPE - 1));
                    const float4 j4 = o[slot], t4 = o[slot + plane];
                    sumJ += V3<T>(j4.x, j4.y, j4.z);
                    sumT += V3<T>(t4.x, t4.y, t4.z);
                }
                b.v += sumJ * invMass;
                b.w += applyIinv(sumT);
            } else {
                for (int i = incBegin; i < incBegin + incCount; ++i) {
                    const int code = M.inc[i];
                    const int c = code & 0xffffff;
                    const bool sideB = (code >> 24) != 0;
                    V3<T> J(sJ[c * 3 + 0], sJ[c * 3 + 1], sJ[c * 3 + 2]);
                    const V3<T> local = sideB ? V3<T>(M.cable(CF_LBX, c), M.cable(CF_LBY, c), M.cable(CF_LBZ, c))
                                              : V3<T>(M.cable(CF_LAX, c), M.cable(CF_LAY, c), M.cable(CF_LAZ, c));
                    if (sideB) J = -J;
                    const V3<T> rel = (b.B * local + b.x) - b.x;
                    b.v += J * invMass;
                    b.w += b.Iinv * rel.cross(J);
                }
            }
        }
        if (TGB_EXCHANGE_UNION) __syncwarp(); /* the next step's contact rows reuse the impulse region */
#ifdef TGB_STATS
        if (a.stats && warpLane == 0) {
            /* cycles of this warp in the body phase, the cable phase and the gather */
            atomicAdd(a.stats + 100, (unsigned long long)(statB - statA));
            atomicAdd(a.stats + 101, (unsigned long long)(statC - statB));
            wCableCyc += (unsigned long long)(statC - statB);
            atomicAdd(a.stats + 102, (unsigned long long)(clock64() - statC));
        }
#endif
    }
#undef CT

#ifdef TGB_STATS
    if (a.stats && warpLane == 0) {
        const unsigned long long cyc = (unsigned long long)(clock64() - statT0);
        atomicAdd(a.stats + 51, cyc); atomicMax(a.stats + 52, cyc); atomicMin(a.stats + 53, cyc); atomicAdd(a.stats + 54, 1ULL);
        const int bucket = (int)min(31ULL, cyc * 16ULL / (unsigned long long)(a.n_steps * 14000LL)); /* cycles per step / 14000 * 16 */
        atomicAdd(a.stats + 64 + bucket, 1ULL);
        const int statCta = blockIdx.x + (a.env_list != nullptr && LPE == 32 ? 2048 : 0); /* tail split: the one-warp-per-environment launch apart */
        if (statCta < 4096 && tid < 32) {
            unsigned long long* r = a.stats + 128 + 8 * statCta;
            atomicAdd(r + 0, cyc); atomicAdd(r + 1, wCoupled); atomicAdd(r + 2, wCoupledCyc); atomicAdd(r + 3, wDetectCyc);
            atomicAdd(r + 4, wMem); atomicAdd(r + 5, wMany); atomicAdd(r + 6, wCableCyc);
        }
    }
#endif
    /* ---- store: bodies back into the record, record back to HBM, frame, status ---------------------------- */
    if (hasBody) {
        T* s = sState + lane * TGB_NBS;
        s[0] = b.x.x; s[1] = b.x.y; s[2] = b.x.z;
        s[3] = b.q.x; s[4] = b.q.y; s[5] = b.q.z; s[6] = b.q.w;
        s[7] = b.v.x; s[8] = b.v.y; s[9] = b.v.z;
        s[10] = b.w.x; s[11] = b.w.y; s[12] = b.w.z;
    }
    /* status: non-finite state, mass centre outside the height window */
    {
        bool bad = false;
        if (hasBody) {
            const T sum = b.x.x + b.x.y + b.x.z + b.q.x + b.q.y + b.q.z + b.q.w + b.v.x + b.v.y + b.v.z + b.w.x + b.w.y + b.w.z;
            bad = !(tabs(sum) <= T(1e30));
        }
        T my = hasBody ? M.body(BF_MASS, bidx) * b.x.y : T(0);
        T mm = hasBody ? M.body(BF_MASS, bidx) : T(0);
        unsigned anyBad = 0;
#pragma unroll
        for (int off = LPE / 2; off > 0; off >>= 1) {
            my += __shfl_xor_sync(fullMask, my, off);
            mm += __shfl_xor_sync(fullMask, mm, off);
        }
        anyBad = __ballot_sync(fullMask, bad);
        const unsigned groupMask = (LPE == 32) ? 0xffffffffu : (((1u << LPE) - 1u) << groupLane0);
        int st = 0;
#pragma unroll
        for (int off = LPE / 2; off > 0; off >>= 1) statusAcc |= __shfl_xor_sync(fullMask, statusAcc, off);
        st |= statusAcc;
        if (anyBad & groupMask) st |= TGB_STATUS_NONFINITE;
        if (a.height_lo <= a.height_hi && mm > T(0)) {
            const double h = Oy + (double)(my / mm);
            if (!(h > a.height_lo && h < a.height_hi)) st |= TGB_STATUS_HEIGHT;
        }
        if (live && lane == 0 && st != 0) a.status[env] |= st;
        if (a.min_clearance != nullptr) {
            float clr = (float)pairClearance;
#pragma unroll
            for (int off = LPE / 2; off > 0; off >>= 1) clr = fminf(clr, __shfl_xor_sync(fullMask, clr, off));
            if (live && lane == 0 && clr < a.min_clearance[env]) a.min_clearance[env] = clr;
        }
        if (a.coupled_steps != nullptr && live && lane == 0 && doWorld)
            a.coupled_steps[env] = (coupledSteps < 0xfffff ? coupledSteps : 0xfffff) | ((coupledRecent < 2047 ? coupledRecent : 2047) << 20);
    }
    if (WIDE) __syncthreads(); else __syncwarp();
    if (live) {
        vec_t* gs = reinterpret_cast<vec_t*>(reinterpret_cast<T*>(a.state) + (size_t)env * a.stride);
        const vec_t* ss = reinterpret_cast<const vec_t*>(sState);
        for (int i = lane; i < a.stride / VN; i += LPE) gs[i] = ss[i];
        if (warmOn && doWorld) {
            T* gw = reinterpret_cast<T*>(a.warm) + (size_t)env * warmWords;
            for (int i = lane; i < warmWords; i += LPE) gw[i] = sWarm[i];
        }
        if (cpgOn) {
            double* cs = a.cpg_state + (size_t)env * TGB_CPG_STATE_DOUBLES(C);
            for (int i = lane; i < 3 * C; i += LPE) { cs[i] = sX[(i % 3) * C + i / 3]; cs[3 * C + 2 + i] = sPid[i]; }
            if (lane == 0) { cs[3 * C] = cpgUpdateTime; cs[3 * C + 1] = cpgCableTime; }
        }
        if (lane == 0) {
            double* f = a.frame + (size_t)env * TGB_NFR;
            f[FR_O] = Ox; f[FR_O + 1] = Oy; f[FR_O + 2] = Oz;
            f[FR_V] = Vx; f[FR_V + 1] = Vy; f[FR_V + 2] = Vz;
            f[FR_CTRLTIME] = ctrlTime;
            if (sineRecurrence && doModels) f[FR_SPARE] = 1.0;
            if (simpleOn && ctrlKind == TGB_CTRL_SERIALIZED_SPINE) f[FR_SPARE] = frameSpare;
        }
        if (sineRecurrence && doModels) {
            double* gsine = a.sine_state + (size_t)env * 48;
            for (int c = lane; c < 24 && c < C; c += LPE) { gsine[2 * c] = sSine[c]; gsine[2 * c + 1] = sSine[24 + c]; }
        }
    }
}

#undef warmOn
#undef warmPairCap
#undef warmWords
#undef sWarm
#undef sWarmPair

/* Shared memory a launch of this kernel needs; mirrors the carve-up above. */
inline size_t tgb_step_smem_bytes(const tgb_step_args& a, int lpe, int threads, size_t sizeofT) {
    const int groupsPerBlock = a.wide ? 32 / lpe : threads / lpe;
    const int modelTWords = (a.model_t_words + 3) & ~3;
    const int jWords = sizeofT == 4 ? a.max_inc * lpe * 8 : 3 * a.C;
    const int ctWords = a.cap_contacts * (a.n_obst > 0 ? (int)TGB_NCT_OBST * a.n_dyn : ((int)TGB_NCT - (a.warm != nullptr && !a.wide ? 0 : 2)) * a.R) +
                        ((a.n_pairs > 0 && a.pair_cap > 0) ? a.pair_cap * (int)TGB_NPR : 0);
    int groupTWords = a.stride + 12 * lpe + (TGB_EXCHANGE_UNION ? (jWords > ctWords ? jWords : ctWords) : jWords + ctWords) + (a.n_pairs > 0 ? a.n_pairs + lpe : 0) +
                      (a.warm != nullptr ? TGB_WARM_STATE_WORDS(a.cap_contacts, a.n_obst > 0 ? a.n_dyn : a.R, (a.n_pairs > 0 && a.pair_cap > 0) ? a.pair_cap : 0) : 0);
    groupTWords = (groupTWords + 3) & ~3;
    if (lpe < 32) { while ((groupTWords & 31) != (lpe < 4 ? 4 : lpe)) groupTWords += 4; }
    const int groupDWords = tgb_group_double_words(a, lpe);
    size_t bytes = (size_t)(modelTWords + groupsPerBlock * groupTWords) * sizeofT;
    bytes = (bytes + 7) & ~(size_t)7;
    bytes += (size_t)groupsPerBlock * groupDWords * sizeof(double);
    bytes += (size_t)((a.model_i_words + 3) & ~3) * sizeof(int);
    if (tgb_cpg_topo_staged(a)) bytes += (size_t)((a.C + 1) + 2 * a.cpg_n_couplings) * sizeof(int);
    return bytes + 16;
}

} // namespace tgb
