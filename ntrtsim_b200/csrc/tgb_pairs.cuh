/*
 * tgb_pairs.cuh -- contact between the bodies of one environment (rod x rod, vertebra x vertebra).
 *
 * Reference behaviour replaced: tgWorldBulletPhysicsImpl.cpp:82-107 builds the world with btAxisSweep3 and
 * btDefaultCollisionConfiguration, i.e. every pair of rigid bodies collides (tgRodInfo.cpp:99-113 btCylinderShape,
 * tgSphereInfo.cpp:100-115 btSphereShape), and the constraint solver couples the two bodies of a manifold point.
 *
 * Manifold (same scheme as the CPU oracle's generateBodyPairContacts): a cylinder counts as the capsule inscribed in it (its axis
 * segment shortened by the radius at both ends, swept by the radius), a sphere as a segment of length zero; one contact at the
 * closest points of the two core segments, two (the ends of the overlap) when the segments are parallel within kParallel and
 * overlap; a point is a contact when the surface distance is within the contact breaking threshold; at most 4 rows per pair of
 * bodies and pair_cap rows per environment, in the order (body A < body B, shape of A, shape of B, point).  Boxes take no part
 * (static boxes are obstacles, see tgb_step.cuh).
 *
 * Mapping.  Detection covers every step: the lanes of an environment's group take the body pairs round robin.  A pair is tested
 * from the published poses (a cull by bounding spheres, then the closest points of the core segments) and the distance found is
 * kept as the pair's slack; every step each body publishes a bound on how far any point of its shapes has moved, the slack shrinks
 * by the two bodies' bounds, and the pair is tested again only when the slack is used up -- exactly conservative, so the result
 * is the same as testing every pair on every step (tgb_step.cuh).  Nothing else happens while no pair is within the contact
 * band.  Otherwise the lanes that found one emit the rows one after the other (row order = pair order, as in the oracle), and the
 * bodies of the warp leave their independent per-lane solvers for this step (solveCoupledWarp): rows against the static world stay
 * with their body's lane, a row between two bodies is worked on by the two lanes that own the bodies -- each forms its side's
 * Jacobian block from its own registers, the two relative velocities meet through one shuffle -- and every iteration of the
 * projected Gauss-Seidel visits the rows in the oracle's order (static-world rows body by body, then the pair rows; normal rows,
 * then friction rows).
 */
#pragma once
#include "tgb_dev.h"
#include "tgb_math.cuh"

namespace tgb {

#ifdef TGB_STATS
/* instrumentation builds: per-section cycles of solveCoupledReg, added to stats[105..] by the first lane of the mask */
#define TGB_PAIR_TICK(slot) do { if (statsP && warpLane == 0) { const long long now_ = clock64(); atomicAdd(statsP + (slot), (unsigned long long)(now_ - tick_)); tick_ = now_; } } while (0)
#else
#define TGB_PAIR_TICK(slot) do { } while (0)
#endif

template <typename T> struct ContactSolveOut { V3<T> dLin, dAng, push, turn; };
/* result of solveCoupledWarp: `solved` = the lane's environment was coupled and `o` is its result; statMem: instrumentation only */
template <typename T> struct CoupledOut { ContactSolveOut<T> o; bool solved; bool statMem; };

template <typename T> TGB_D T pairBreakingThreshold() { return T(0.02); }
template <typename T> TGB_D T pairParallel() { return T(1e-4); }

template <typename T>
struct CoreSegment {
    V3<T> p, d; /* centre, unit direction */
    T h, r;     /* half length of the core segment, radius of the sweep */
};

template <typename T> TGB_D T clampT(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* core segment of collision shape si of the body whose published pose record (pos3, basis9) is `pose`; false for a box */
template <typename T, typename Model>
TGB_D bool coreSegment(const Model& M, const T* pose, int si, CoreSegment<T>& g) {
    if (M.shape(SF_KIND, si) == T(TGB_SHAPE_BOX)) return false;
    const V3<T> x(pose[0], pose[1], pose[2]);
    M3<T> B;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) B.m[i][j] = pose[3 + i * 3 + j];
    const V3<T> spos(M.shape(SF_POSX, si), M.shape(SF_POSY, si), M.shape(SF_POSZ, si));
    const V3<T> r1(M.shape(SF_ROT0 + 1, si), M.shape(SF_ROT0 + 4, si), M.shape(SF_ROT0 + 7, si));
    g.p = B * spos + x;
    g.d = B * r1;
    g.r = M.shape(SF_RADIUS, si);
    g.h = M.shape(SF_SEGH, si);
    return true;
}

/* the same for a body that is one capsule in the body frame (BF_SEGH >= 0): centre = origin, direction = the basis' second column --
   bit for bit what coreSegment gives for such a shape */
template <typename T, typename Model>
TGB_D void coreSegmentOfRod(const Model& M, const T* pose, int body, CoreSegment<T>& g) {
    g.p = V3<T>(pose[0], pose[1], pose[2]);
    g.d = V3<T>(pose[4], pose[7], pose[10]);
    g.r = M.body(BF_SEGR, body);
    g.h = M.body(BF_SEGH, body);
}

/* closest points A(s) = pa + s da, B(t) = pb + t db of two segments given by centre, unit direction and half length */
template <typename T>
TGB_D void closestSegmentParams(const CoreSegment<T>& A, const CoreSegment<T>& B, T& sOut, T& tOut, T& bOut, T& cOut, T& fOut) {
    const V3<T> r = A.p - B.p;
    const T b = A.d.dot(B.d), c = A.d.dot(r), f = B.d.dot(r);
    const T denom = T(1) - b * b;
    T s = T(0), t = T(0);
    if (A.h > T(0) && denom > T(1e-12)) s = clampT(tdiv(b * f - c, denom), -A.h, A.h); /* double: exact division, the oracle's */
    t = b * s + f;
    if (t < -B.h) { t = -B.h; s = clampT(b * t - c, -A.h, A.h); }
    else if (t > B.h) { t = B.h; s = clampT(b * t - c, -A.h, A.h); }
    sOut = s; tOut = t; bOut = b; cOut = c; fOut = f;
}

/* Detection: the smallest surface distance `dist` between the capsules of bodies (ia, ib) (a lower bound from the bounding spheres
   when those are apart); returns whether a pair of shapes is within the contact band (1e-3 of slack for round-off: emitPairRows
   applies the exact test).  Also folds the distance into the running minimum `clearance`. */
/* A contact found by the detection itself (two plain rods / balls, one point): what emitCapsulePair would compute again.  valid = false:
   the pair takes emitPairRows (compound bodies, the two-point manifold of parallel rods, axes that meet). */
template <typename T>
struct PairRecord {
    bool valid;
    V3<T> n, pA, pB;
    T dist;
};

template <typename T, typename Model>
TGB_D bool pairNear(const Model& M, const T* sPose, int ia, int ib, T& clearance, T& dist, PairRecord<T>& rec, bool* deferCompound = nullptr) {
    rec.valid = false;
    const T* pa = sPose + 12 * ia;
    const T* pb = sPose + 12 * ib;
    {
        /* bounding spheres about the body origins */
        const V3<T> d(pa[0] - pb[0], pa[1] - pb[1], pa[2] - pb[2]);
        const T reach = M.body(BF_BOUND, ia) + M.body(BF_BOUND, ib);
        const T far = reach + pairBreakingThreshold<T>() + T(1e-3);
        const T d2 = d.length2();
        if (d2 > far * far) { dist = tsqrtNoBranch(d2) - reach; return false; }
    }
    if (M.body(BF_SEGH, ia) >= T(0) && M.body(BF_SEGH, ib) >= T(0)) {
        /* two plain rods / balls: straight-line code */
        CoreSegment<T> ga, gb;
        coreSegmentOfRod<T>(M, pa, ia, ga);
        coreSegmentOfRod<T>(M, pb, ib, gb);
        T s, t, b, c, f;
        closestSegmentParams<T>(ga, gb, s, t, b, c, f);
        const V3<T> ca = ga.p + ga.d * s, cb = gb.p + gb.d * t;
        const V3<T> dv = ca - cb;
        const T d = tsqrtNoBranch(dv.length2()); /* double: sqrt, as in emitCapsulePair */
        const T dd = d - ga.r - gb.r;
        clearance = dd < clearance ? dd : clearance;
        dist = dd;
        const bool near = dd <= pairBreakingThreshold<T>() + T(1e-3);
        if (near && d > T(1e-9) && !(ga.h > T(0) && gb.h > T(0) && T(1) - b * b < pairParallel<T>())) {
            /* the one contact of this pair, with emitCapsulePair's operations */
            rec.valid = true;
            rec.n = dv * trcp(d);
            rec.pA = ca - rec.n * ga.r; rec.pB = cb + rec.n * gb.r;
            rec.dist = dd;
        }
        return near;
    }
    if (deferCompound != nullptr) {
        /* bodies of several shapes: the caller spreads the pairs of shapes over the lanes of the group (pairNearCompoundGroup) */
        *deferCompound = true;
        dist = T(0);
        return false;
    }
    bool near = false;
    T best = T(1e30);
    const int sa0 = M.bodyi(BI_SHAPE_BEGIN, ia), sa1 = sa0 + M.bodyi(BI_SHAPE_COUNT, ia);
    const int sb0 = M.bodyi(BI_SHAPE_BEGIN, ib), sb1 = sb0 + M.bodyi(BI_SHAPE_COUNT, ib);
    for (int sa = sa0; sa < sa1; ++sa) {
        CoreSegment<T> ga;
        if (!coreSegment<T>(M, pa, sa, ga)) continue;
        for (int sb = sb0; sb < sb1; ++sb) {
            CoreSegment<T> gb;
            if (!coreSegment<T>(M, pb, sb, gb)) continue;
            T s, t, b, c, f;
            closestSegmentParams<T>(ga, gb, s, t, b, c, f);
            const V3<T> dv = (ga.p + ga.d * s) - (gb.p + gb.d * t);
            const T dd = tsqrtNoBranch(dv.length2()) - ga.r - gb.r;
            best = dd < best ? dd : best;
            near = near || dd <= pairBreakingThreshold<T>() + T(1e-3);
        }
    }
    clearance = best < clearance ? best : clearance;
    dist = best;
    return near;
}

/* The same test for two bodies made of several shapes, by all lanes of the environment's group together: the pairs of shapes go round
   the lanes (a TetraSpine vertebra has six rods: 36 segment-segment tests per pair of bodies, which one lane used to walk alone while
   the others waited -- the population evaluation lost half its rate to it), then the smallest distance and the near flag are reduced
   over the group.  Same distances as pairNear's loop; the minimum of a set does not depend on the order. */
template <typename T, typename Model, int LPE>
TGB_D void pairNearCompoundGroup(const Model& M, const T* sPose, int ia, int ib, int lane, unsigned groupMask, T& bestOut, T& exactOut, bool& nearOut) {
    const T* pa = sPose + 12 * ia;
    const T* pb = sPose + 12 * ib;
    const int sa0 = M.bodyi(BI_SHAPE_BEGIN, ia), na = M.bodyi(BI_SHAPE_COUNT, ia);
    const int sb0 = M.bodyi(BI_SHAPE_BEGIN, ib), nb = M.bodyi(BI_SHAPE_COUNT, ib);
    T best = T(1e30), exact = T(1e30);
    bool near = false;
    for (int t = lane; t < na * nb; t += LPE) {
        const int sa = sa0 + t / nb, sb = sb0 + t % nb;
        CoreSegment<T> ga, gb;
        if (!coreSegment<T>(M, pa, sa, ga) || !coreSegment<T>(M, pb, sb, gb)) continue;
        /* (no bounding test per pair of shapes: a lower bound from the spheres about two long rods is far below their distance and
           would use the pair's slack up at once -- measured: the pairs were then tested on every step, 2.3e7 -> 1.4e7 env-steps/s) */
        T s, tt, b, c, f;
        closestSegmentParams<T>(ga, gb, s, tt, b, c, f);
        const V3<T> dv = (ga.p + ga.d * s) - (gb.p + gb.d * tt);
        const T dd = tsqrtNoBranch(dv.length2()) - ga.r - gb.r;
        best = dd < best ? dd : best;
        exact = dd < exact ? dd : exact;
        near = near || dd <= pairBreakingThreshold<T>() + T(1e-3);
    }
#pragma unroll
    for (int off = LPE / 2; off > 0; off >>= 1) {
        const T ob = __shfl_xor_sync(groupMask, best, off), oe = __shfl_xor_sync(groupMask, exact, off);
        best = ob < best ? ob : best;
        exact = oe < exact ? oe : exact;
    }
    nearOut = (__ballot_sync(groupMask, near) & groupMask) != 0u;
    bestOut = best; exactOut = exact;
}

/* The contacts of one pair of capsules (bodies ia < ib) as rows base + count, ... of the environment's pair rows; `count` counts the rows
   of this pair of bodies (at most 4).  Mirrors the oracle's generateBodyPairContacts expression by expression. */
/* where emission finds the persistent manifold of the environment's pair rows (tgb_dev.h TGB_WARM_*); entries == nullptr: no warm start */
template <typename T>
struct PairWarm {
    const T* entries; /* [cap][TGB_WARM_WORDS] */
    int cap;
    T factor;
};

template <typename T>
TGB_D void emitCapsulePair(const CoreSegment<T>& ga, const CoreSegment<T>& gb, int ia, int ib, int saLocal, int sbLocal, T mu, T restitution,
                           const PairWarm<T>& warm, T* rows, int base, int cap, int& count) {
    T s0, t0, b, c, f;
    closestSegmentParams<T>(ga, gb, s0, t0, b, c, f);
    T sp[2] = {s0, s0}, tp[2] = {t0, t0};
    int np = 1;
    if (ga.h > T(0) && gb.h > T(0) && T(1) - b * b < pairParallel<T>()) {
        /* B's end points in A's axis parameter: s(t) = t b - c */
        const T e0 = -gb.h * b - c, e1 = gb.h * b - c;
        const T emin = e0 < e1 ? e0 : e1, emax = e0 < e1 ? e1 : e0;
        const T lo = -ga.h > emin ? -ga.h : emin, hi = ga.h < emax ? ga.h : emax;
        if (hi - lo > T(1e-9)) {
            np = 2;
            sp[0] = lo; tp[0] = clampT(b * lo + f, -gb.h, gb.h);
            sp[1] = hi; tp[1] = clampT(b * hi + f, -gb.h, gb.h);
        }
    }
    for (int k = 0; k < np; ++k) {
        const V3<T> ca = ga.p + ga.d * sp[k], cb = gb.p + gb.d * tp[k];
        const V3<T> dv = ca - cb;
        const T d = tsqrt(dv.length2());
        const T dist = d - ga.r - gb.r;
        if (dist > pairBreakingThreshold<T>()) continue;
        if (count >= 4 || base + count >= cap) continue;
        V3<T> n;
        if (d > T(1e-9)) n = dv * trcp(d); /* double: 1 / d, the oracle's */
        else {
            /* the axes meet: any direction perpendicular to both */
            const V3<T> x = ga.d.cross(gb.d);
            const T xl = x.length();
            if (xl > T(1e-9)) n = x * (T(1) / xl);
            else {
                /* btPlaneSpace1 of A's axis */
                const V3<T> a = ga.d;
                if (tabs(a.z) > T(0.7071067811865475244008443621048490)) {
                    const T kk = T(1) / tsqrt(a.y * a.y + a.z * a.z);
                    n = V3<T>(T(0), -a.z * kk, a.y * kk);
                } else {
                    const T kk = T(1) / tsqrt(a.x * a.x + a.y * a.y);
                    n = V3<T>(-a.y * kk, a.x * kk, T(0));
                }
            }
        }
        const V3<T> pA = ca - n * ga.r, pB = cb + n * gb.r; /* the contact points on the two surfaces */
        T* row = rows + (base + count) * TGB_NPR;
        row[PR_N + 0] = n.x; row[PR_N + 1] = n.y; row[PR_N + 2] = n.z;
        row[PR_RA + 0] = pA.x; row[PR_RA + 1] = pA.y; row[PR_RA + 2] = pA.z;
        row[PR_RB + 0] = pB.x; row[PR_RB + 1] = pB.y; row[PR_RB + 2] = pB.z;
        row[PR_LAMN] = dist; row[PR_LAMPUSH] = restitution;
        row[PR_MU] = mu; row[PR_IDS] = (T)(ia + 32 * ib);
        /* the point's impulses at the end of the last step, scaled (0 when the point is new) */
        const T key = (T)((((ia + 32 * ib) * 64 + saLocal) * 64 + sbLocal) * 2 + k);
        T wN = T(0), wT = T(0);
        if (warm.entries != nullptr)
            for (int j = 0; j < warm.cap; ++j) {
                const T* e = warm.entries + j * TGB_WARM_WORDS;
                if (e[0] == key) { wN = e[1] * warm.factor; wT = e[2] * warm.factor; }
            }
        row[PR_KEY] = key; row[PR_INVDIAGT] = wN; row[PR_RHST] = wT;
        ++count;
    }
}

/* Emission of the contact the detection already holds (PairRecord): row `base` of the environment's pair rows; returns the rows written */
template <typename T, typename Model>
TGB_D int emitRecordedPair(const Model& M, const PairRecord<T>& rec, int ia, int ib, const PairWarm<T>& warm, T* rows, int base, int cap) {
    if (rec.dist > pairBreakingThreshold<T>() || base >= cap) return 0;
    T mu = M.body(BF_FRICTION, ia) * M.body(BF_FRICTION, ib);
    mu = mu < T(-10) ? T(-10) : (mu > T(10) ? T(10) : mu);
    T* row = rows + base * TGB_NPR;
    row[PR_N + 0] = rec.n.x; row[PR_N + 1] = rec.n.y; row[PR_N + 2] = rec.n.z;
    row[PR_RA + 0] = rec.pA.x; row[PR_RA + 1] = rec.pA.y; row[PR_RA + 2] = rec.pA.z;
    row[PR_RB + 0] = rec.pB.x; row[PR_RB + 1] = rec.pB.y; row[PR_RB + 2] = rec.pB.z;
    row[PR_LAMN] = rec.dist; row[PR_LAMPUSH] = M.body(BF_RESTITUTION, ia) * M.body(BF_RESTITUTION, ib);
    row[PR_MU] = mu; row[PR_IDS] = (T)(ia + 32 * ib);
    const T key = (T)((((ia + 32 * ib) * 64) * 64) * 2);
    T wN = T(0), wT = T(0);
    if (warm.entries != nullptr)
        for (int j = 0; j < warm.cap; ++j) {
            const T* e = warm.entries + j * TGB_WARM_WORDS;
            if (e[0] == key) { wN = e[1] * warm.factor; wT = e[2] * warm.factor; }
        }
    row[PR_KEY] = key; row[PR_INVDIAGT] = wN; row[PR_RHST] = wT;
    return 1;
}

/* Emission for bodies made of several shapes (compound bodies); out of line. */
template <typename T, typename Model>
__device__ __noinline__ int emitCompoundPairRows(const Model M, const T* sPose, int ia, int ib, T mu, T restitution, const PairWarm<T> warm, T* rows,
                                                 int base, int cap) {
    const T* pa = sPose + 12 * ia;
    const T* pb = sPose + 12 * ib;
    int count = 0;
    const int sa0 = M.bodyi(BI_SHAPE_BEGIN, ia), sa1 = sa0 + M.bodyi(BI_SHAPE_COUNT, ia);
    const int sb0 = M.bodyi(BI_SHAPE_BEGIN, ib), sb1 = sb0 + M.bodyi(BI_SHAPE_COUNT, ib);
    for (int sa = sa0; sa < sa1; ++sa) {
        CoreSegment<T> ga;
        if (!coreSegment<T>(M, pa, sa, ga)) continue;
        for (int sb = sb0; sb < sb1; ++sb) {
            CoreSegment<T> gb;
            if (!coreSegment<T>(M, pb, sb, gb)) continue;
            emitCapsulePair<T>(ga, gb, ia, ib, sa - sa0, sb - sb0, mu, restitution, warm, rows, base, cap, count);
        }
    }
    return count;
}

/* Emission: the contacts of bodies (ia, ib) as rows base, base + 1, ... of the environment's pair rows; returns the number written.
   Two plain rods / balls in line, compound bodies out of line. */
template <typename T, typename Model>
TGB_D int emitPairRows(const Model& M, const T* sPose, int ia, int ib, const PairWarm<T>& warm, T* rows, int base, int cap) {
    T mu = M.body(BF_FRICTION, ia) * M.body(BF_FRICTION, ib);
    mu = mu < T(-10) ? T(-10) : (mu > T(10) ? T(10) : mu);
    const T restitution = M.body(BF_RESTITUTION, ia) * M.body(BF_RESTITUTION, ib);
    if (M.body(BF_SEGH, ia) >= T(0) && M.body(BF_SEGH, ib) >= T(0)) {
        CoreSegment<T> ga, gb;
        coreSegmentOfRod<T>(M, sPose + 12 * ia, ia, ga);
        coreSegmentOfRod<T>(M, sPose + 12 * ib, ib, gb);
        int count = 0;
        emitCapsulePair<T>(ga, gb, ia, ib, 0, 0, mu, restitution, warm, rows, base, cap, count);
        return count;
    }
    return emitCompoundPairRows<T>(M, sPose, ia, ib, mu, restitution, warm, rows, base, cap);
}

/* World inverse inertia as the coupled solvers apply it: double keeps the matrix (the oracle's operations), single precision the factored
   form B (invI o (B^T v)) of the step kernel -- the nine words of the matrix are not held in registers there. */
template <typename T>
struct InvInertia {
    M3<T> B;      /* basis (single precision) or the world inverse inertia matrix itself (double) */
    V3<T> invI;   /* local inverse inertia (single precision) */
    bool factored;
    TGB_D V3<T> operator*(const V3<T>& v) const {
        if (factored) {
            const V3<T> l = B.tmul(v);
            return B * V3<T>(invI.x * l.x, invI.y * l.y, invI.z * l.z);
        }
        return B * v;
    }
};

/* What the coupled solvers need of the lane's body (they are inlined into the step kernel: these are the kernel's registers).  A lane
   without a dynamic body of its own takes part with zero inverse mass and inertia. */
template <typename T>
struct CoupledBody {
    V3<T> pos, gnv;
    InvInertia<T> inertia;
    T invMassV, muV;
    bool mine;
    TGB_D V3<T> x() const { return pos; }
    TGB_D T invMass() const { return invMassV; }
    TGB_D T mu() const { return muV; }
    TGB_D V3<T> normal() const { return gnv; }
    TGB_D InvInertia<T> Iinv() const { return inertia; }
};

/* solveCoupledReg is compiled for (own contact points per lane, pair rows per environment) = (1, 1) -- a rod end on the ground, one
   rod-rod contact: the commonest case by far -- and (2, 2); environments beyond that take solveCoupledMem.  A single warp issues
   roughly one instruction every four cycles here (dependent chains), so the executed instruction count of a coupled step is what
   sets the duration of a launch: the (1, 1) form runs 4 local rows instead of 8. */
#ifndef TGB_PAIR_REG_ROWS
#define TGB_PAIR_REG_ROWS 2
#endif

/* The contact solve of the environments of a warp that have rows between two bodies ("coupled" environments).  Three steps, each
   entered together by all lanes of those environments (`mask`; lanes without a body and static bodies included), so that the coupled
   environments of a warp go through the iterations side by side instead of one after the other -- a launch ends with its slowest
   warp, and rods that lean on each other do so for thousands of steps:
     setupPairRows     the two-body rows' constants (setupContactConstraint / setupFrictionConstraint with two bodies)
     solveCoupledReg   the iterations with everything in registers: at most two own rows per lane, TGB_PAIR_REG_ROWS pair rows per environment
     solveCoupledMem   the same iterations over the rows in shared memory, any counts
   Which of the two solvers an environment takes depends on its own rows only, and both perform the same operations in the same
   order; an environment without pair rows keeps its per-lane solver (tgb_step.cuh) whatever its warp neighbours do.  Results
   therefore do not depend on which environments share a warp.

   Order of a Gauss-Seidel iteration = the oracle's: normal rows against the static world (independent between bodies, so every lane
   does its own at the same time), the pair rows' normals in row order, then the friction rows the same way.  A pair row is worked on
   by the two lanes that own its bodies: each forms its side J_side . dV of the relative velocity from its registers, one shuffle
   brings the other side over, both lanes compute the same impulse and apply it to their own body. */
template <typename T, bool kFast>
__device__ __forceinline__ void setupPairRows(unsigned mask, T* pr, int nPairRows, int lane, int groupLane0, const CoupledBody<T>& body, V3<T> vel,
                                           V3<T> w, V3<T> extLin, V3<T> extAng, T dt) {
    const V3<T> x = body.x();
    const InvInertia<T> Iinv = body.Iinv();
    const T invMass = body.invMass();
    const int warpLane = groupLane0 + lane;
    const int maxP = __reduce_max_sync(mask, nPairRows);
    for (int p = 0; p < maxP; ++p) {
        const bool have = p < nPairRows;
        T* row = pr + (have ? p : 0) * TGB_NPR;
        const int ids = (int)row[PR_IDS];
        const int ia = ids & 31, ib = ids >> 5;
        const bool isA = have && lane == ia, isB = have && lane == ib;
        const int src = (isA || isB) ? groupLane0 + (isA ? ib : ia) : warpLane;
        const V3<T> n(row[PR_N + 0], row[PR_N + 1], row[PR_N + 2]);
        const int po = isB ? PR_RB : PR_RA;
        const V3<T> r = V3<T>(row[po + 0], row[po + 1], row[po + 2]) - x;
        const T dist = row[PR_LAMN], combinedRestitution = row[PR_LAMPUSH];
        const T warmN = row[PR_INVDIAGT], warmT = row[PR_RHST];
        const V3<T> rxn = r.cross(n);
        const V3<T> angN = Iinv * rxn;
        const T dN = invMass + n.dot(angN.cross(r));
        const V3<T> vp = vel + w.cross(r);
        const T vdn = n.dot(vel + extLin) + rxn.dot(w + extAng);
        const T o_dN = __shfl_sync(mask, dN, src), o_vdn = __shfl_sync(mask, vdn, src);
        const V3<T> o_vp(__shfl_sync(mask, vp.x, src), __shfl_sync(mask, vp.y, src), __shfl_sync(mask, vp.z, src));
        const T denom = (isB ? o_dN : dN) + (isB ? dN : o_dN);
        const T invDiagN = kFast ? trcp(denom) : T(1) / denom;
        const V3<T> v = (isB ? o_vp : vp) - (isB ? vp : o_vp);
        const T relVel = n.dot(v);
        T restitution = combinedRestitution * -relVel;
        if (restitution <= T(0)) restitution = T(0);
        const V3<T> lat = v - n * relVel;
        const T lat2 = lat.length2();
        V3<T> t;
        if (lat2 > (kFast ? T(1.1920928955078125e-7) : T(2.2204460492503131e-16))) t = lat * (kFast ? trsqrt(lat2) : T(1) / tsqrt(lat2));
        else if (tabs(n.z) > T(0.7071067811865475244008443621048490)) {
            const T kk = T(1) / tsqrt(n.y * n.y + n.z * n.z);
            t = V3<T>(T(0), -n.z * kk, n.y * kk);
        } else {
            const T kk = T(1) / tsqrt(n.x * n.x + n.y * n.y);
            t = V3<T>(-n.y * kk, n.x * kk, T(0));
        }
        const T velDotN = (isB ? o_vdn : vdn) - (isB ? vdn : o_vdn);
        T velocityError = restitution - velDotN;
        const T penetration = dist;
        T erp = T(0.8);
        if (penetration > T(-0.04)) erp = T(0.2);
        T positionalError = T(0);
        if (penetration > T(0)) velocityError -= tdiv(penetration, dt);
        else positionalError = tdiv(-penetration * erp, dt);
        const T penetrationImpulse = positionalError * invDiagN;
        const T velocityImpulse = velocityError * invDiagN;
        T rhsN, rhsPen;
        if (penetration > T(-0.04)) { rhsN = penetrationImpulse + velocityImpulse; rhsPen = T(0); }
        else { rhsN = velocityImpulse; rhsPen = penetrationImpulse; }
        const V3<T> rxt = r.cross(t);
        const V3<T> angT = Iinv * rxt;
        const T dT = invMass + t.dot(angT.cross(r));
        const T rT = t.dot(vel + extLin) + rxt.dot(w);
        const T o_dT = __shfl_sync(mask, dT, src), o_rT = __shfl_sync(mask, rT, src);
        const T denomT = (isB ? o_dT : dT) + (isB ? dT : o_dT);
        const T relT = (isB ? o_rT : rT) - (isB ? rT : o_rT);
        const T invDiagT = kFast ? trcp(denomT) : T(1) / denomT;
        const T rhsT = (T(0) - relT) * invDiagT;
        __syncwarp(mask); /* every lane has read the detection's scratch fields of this row */
        if (isA) {
            row[PR_RA + 0] = r.x; row[PR_RA + 1] = r.y; row[PR_RA + 2] = r.z;
            row[PR_T + 0] = t.x; row[PR_T + 1] = t.y; row[PR_T + 2] = t.z;
            row[PR_INVDIAGN] = invDiagN; row[PR_RHSN] = rhsN; row[PR_RHSPEN] = rhsPen; row[PR_LAMN] = warmN; row[PR_LAMPUSH] = T(0);
            row[PR_INVDIAGT] = invDiagT; row[PR_RHST] = rhsT; row[PR_LAMT] = warmT;
        }
        if (isB) { row[PR_RB + 0] = r.x; row[PR_RB + 1] = r.y; row[PR_RB + 2] = r.z; }
    }
    __syncwarp(mask);
}

/* does this lane's environment have a split-impulse row (group-uniform when every lane of the group passes its own flag)? */
template <typename T>
TGB_D bool groupHasPenetration(unsigned groupMask, bool anyPenOwn, const T* pr, int nPairRows) {
    bool pen = __ballot_sync(groupMask, anyPenOwn) != 0u;
    for (int p = 0; p < nPairRows; ++p) pen = pen || pr[p * TGB_NPR + PR_RHSPEN] != T(0);
    return pen;
}

#define TGB_CTG(slot, f) ct[(slot) * rowStride + (f)]

/* Set-up and iterations in registers: the pair rows come straight from the detection's records (emitPairRows), nothing is written
   back.  Double precision iterates in velocity space in the oracle's order of operations (the parity instrument).  Single precision
   iterates in impulse space: every lane carries, for each of its rows, the part of the row's relative velocity that its own body
   contributes (dv_i = J_i . dV_body) and updates all of them by one multiply-add per applied impulse with the row couplings
   W_ij = J_i M^-1 J_j^T formed once -- the dependent chain of a row visit is 5 operations instead of ~12, and a pair row adds one
   shuffle (the two sides' parts of its relative velocity).  Equal to the velocity-space iteration in exact arithmetic. */
template <typename T, bool kFast, bool OBST, int NC, int PM>
__device__ __forceinline__ ContactSolveOut<T> solveCoupledReg(T* ct, int rowStride, int nc, bool anyPenOwn, const T* pr, T* warmPair, int nPairRows,
                                                            int lane, int groupLane0, const CoupledBody<T>& body, V3<T> vel, V3<T> w, V3<T> extLin,
                                                            V3<T> extAng, T dt, int iterations
#ifdef TGB_STATS
                                                            , unsigned long long* statsP
#endif
                                                            ) {
    constexpr int PMAX = PM;
    /* All 32 lanes of the warp are here (lanes of environments that take another solver pass nc = 0, nPairRows = 0 and ignore the
       result): with the constant full mask a shuffle is one instruction, without the convergence check a run-time mask costs. */
    constexpr unsigned mask = 0xffffffffu;
    const int warpLane = groupLane0 + lane;
    const V3<T> x = body.x(), gn = body.normal();
    const InvInertia<T> Iinv = body.Iinv();
    const T invMass = body.invMass(), muGround = body.mu();
#ifdef TGB_STATS
    long long tick_ = clock64();
#endif
    const int maxP = __reduce_max_sync(mask, nPairRows);
    /* ---- own rows (up to two contact points against the static world), as the step kernel set them up ---- */
    V3<T> on[NC], orxn[NC], oangN[NC], ot[NC], orxt[NC], oangT[NC];
    T oInvN[NC], oRhsN[NC], oRhsPen[NC], oInvT[NC], oRhsT[NC], oMu[NC], oWarmN[NC], oWarmT[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int cc = c < nc ? c : 0; /* a lane without that row reads row 0 (or stale words) and never uses the result */
        on[c] = OBST ? V3<T>(TGB_CTG(cc, CT_NRM + 0), TGB_CTG(cc, CT_NRM + 1), TGB_CTG(cc, CT_NRM + 2)) : gn;
        orxn[c] = V3<T>(TGB_CTG(cc, CT_RXN + 0), TGB_CTG(cc, CT_RXN + 1), TGB_CTG(cc, CT_RXN + 2));
        oangN[c] = V3<T>(TGB_CTG(cc, CT_ANGN + 0), TGB_CTG(cc, CT_ANGN + 1), TGB_CTG(cc, CT_ANGN + 2));
        ot[c] = V3<T>(TGB_CTG(cc, CT_T + 0), TGB_CTG(cc, CT_T + 1), TGB_CTG(cc, CT_T + 2));
        orxt[c] = V3<T>(TGB_CTG(cc, CT_RXT + 0), TGB_CTG(cc, CT_RXT + 1), TGB_CTG(cc, CT_RXT + 2));
        oangT[c] = V3<T>(TGB_CTG(cc, CT_ANGT + 0), TGB_CTG(cc, CT_ANGT + 1), TGB_CTG(cc, CT_ANGT + 2));
        oInvN[c] = TGB_CTG(cc, CT_INVDIAGN); oRhsN[c] = TGB_CTG(cc, CT_RHSN); oRhsPen[c] = TGB_CTG(cc, CT_RHSPEN);
        oInvT[c] = TGB_CTG(cc, CT_INVDIAGT); oRhsT[c] = TGB_CTG(cc, CT_RHST);
        oMu[c] = OBST ? TGB_CTG(cc, CT_MU) : muGround;
        oWarmN[c] = c < nc ? TGB_CTG(cc, CT_LAMN) : T(0); oWarmT[c] = c < nc ? TGB_CTG(cc, CT_LAMT) : T(0);
    }
    TGB_PAIR_TICK(105);
    /* ---- this lane's side of its environment's pair rows (setupContactConstraint / setupFrictionConstraint with two bodies):
       sign +1 (body A), -1 (body B) or 0 (not its row) ---- */
    V3<T> pn[PMAX], prxn[PMAX], pangN[PMAX], pt[PMAX], prxt[PMAX], pangT[PMAX];
    T pInvN[PMAX], pRhsN[PMAX], pRhsPen[PMAX], pInvT[PMAX], pRhsT[PMAX], pMu[PMAX], pSign[PMAX];
    T pWarmN[PMAX], pWarmT[PMAX], pKey[PMAX]; /* warm starting: the point's scaled impulses of the last step, its identity */
    int pSrc[PMAX];
#pragma unroll
    for (int p = 0; p < PMAX; ++p) {
        const bool have = p < nPairRows;
        const T* row = pr + (have ? p : 0) * TGB_NPR;
        const int ids = (int)row[PR_IDS];
        const int ia = ids & 31, ib = ids >> 5;
        const bool isA = have && lane == ia, isB = have && lane == ib;
        const int src = (isA || isB) ? groupLane0 + (isA ? ib : ia) : warpLane;
        pSrc[p] = src;
        pSign[p] = isA ? T(1) : (isB ? T(-1) : T(0));
        const V3<T> n(row[PR_N + 0], row[PR_N + 1], row[PR_N + 2]);
        const int po = isB ? PR_RB : PR_RA;
        const V3<T> r = V3<T>(row[po + 0], row[po + 1], row[po + 2]) - x;
        const T dist = row[PR_LAMN], combinedRestitution = row[PR_LAMPUSH];
        pMu[p] = row[PR_MU];
        pWarmN[p] = (isA || isB) ? row[PR_INVDIAGT] : T(0); pWarmT[p] = (isA || isB) ? row[PR_RHST] : T(0);
        pKey[p] = row[PR_KEY];
        const V3<T> rxn = r.cross(n);
        const V3<T> angN = Iinv * rxn;
        const T dN = invMass + n.dot(angN.cross(r));
        const V3<T> vp = vel + w.cross(r);
        const T vdn = n.dot(vel + extLin) + rxn.dot(w + extAng);
        T o_dN = dN, o_vdn = vdn;
        V3<T> o_vp = vp;
        if (PM == 1 || p < maxP) { /* uniform over the lanes of `mask` */
            o_dN = __shfl_sync(mask, dN, src); o_vdn = __shfl_sync(mask, vdn, src);
            o_vp = V3<T>(__shfl_sync(mask, vp.x, src), __shfl_sync(mask, vp.y, src), __shfl_sync(mask, vp.z, src));
        }
        const T denom = (isB ? o_dN : dN) + (isB ? dN : o_dN);
        const T invDiagN = kFast ? trcp(denom) : T(1) / denom;
        const V3<T> v = (isB ? o_vp : vp) - (isB ? vp : o_vp);
        const T relVel = n.dot(v);
        T restitution = combinedRestitution * -relVel;
        if (restitution <= T(0)) restitution = T(0);
        const V3<T> lat = v - n * relVel;
        const T lat2 = lat.length2();
        V3<T> t;
        if (lat2 > (kFast ? T(1.1920928955078125e-7) : T(2.2204460492503131e-16))) t = lat * (kFast ? trsqrt(lat2) : T(1) / tsqrt(lat2));
        else if (tabs(n.z) > T(0.7071067811865475244008443621048490)) {
            const T kk = T(1) / tsqrt(n.y * n.y + n.z * n.z);
            t = V3<T>(T(0), -n.z * kk, n.y * kk);
        } else {
            const T kk = T(1) / tsqrt(n.x * n.x + n.y * n.y);
            t = V3<T>(-n.y * kk, n.x * kk, T(0));
        }
        const T velDotN = (isB ? o_vdn : vdn) - (isB ? vdn : o_vdn);
        T velocityError = restitution - velDotN;
        const T penetration = dist;
        T erp = T(0.8);
        if (penetration > T(-0.04)) erp = T(0.2);
        T positionalError = T(0);
        if (penetration > T(0)) velocityError -= tdiv(penetration, dt);
        else positionalError = tdiv(-penetration * erp, dt);
        const T penetrationImpulse = positionalError * invDiagN;
        const T velocityImpulse = velocityError * invDiagN;
        T rhsN, rhsPen;
        if (penetration > T(-0.04)) { rhsN = penetrationImpulse + velocityImpulse; rhsPen = T(0); }
        else { rhsN = velocityImpulse; rhsPen = penetrationImpulse; }
        const V3<T> rxt = r.cross(t);
        const V3<T> angT = Iinv * rxt;
        const T dT = invMass + t.dot(angT.cross(r));
        const T rT = t.dot(vel + extLin) + rxt.dot(w);
        T o_dT = dT, o_rT = rT;
        if (PM == 1 || p < maxP) { o_dT = __shfl_sync(mask, dT, src); o_rT = __shfl_sync(mask, rT, src); }
        const T denomT = (isB ? o_dT : dT) + (isB ? dT : o_dT);
        const T relT = (isB ? o_rT : rT) - (isB ? rT : o_rT);
        const T invDiagT = kFast ? trcp(denomT) : T(1) / denomT;
        pn[p] = n; prxn[p] = rxn; pangN[p] = angN; pt[p] = t; prxt[p] = rxt; pangT[p] = angT;
        pInvN[p] = invDiagN; pRhsN[p] = rhsN; pRhsPen[p] = (isA || isB) ? rhsPen : T(0);
        pInvT[p] = invDiagT; pRhsT[p] = (T(0) - relT) * invDiagT;
    }
    bool penHere = anyPenOwn;
#pragma unroll
    for (int p = 0; p < PMAX; ++p) penHere = penHere || pRhsPen[p] != T(0);
    const bool anyPen = __any_sync(mask, penHere);
    TGB_PAIR_TICK(106);

    V3<T> dLin(T(0), T(0), T(0)), dAng(T(0), T(0), T(0)), push(T(0), T(0), T(0)), turn(T(0), T(0), T(0));
    if constexpr (kFast) {
        /* ---- impulse space.  Local rows: own normals, own frictions, pair normals (from PN), pair frictions (from PT).  The side's
           sign is folded into the Jacobian (J = sign (d, r x d)), so a pair row's relative velocity is the SUM of its two sides' parts. */
        constexpr int K = 2 * NC + 2 * PMAX;
        constexpr int PN = 2 * NC, PT = 2 * NC + PMAX; /* first pair normal / pair friction row */
        V3<T> jl[K], ja[K], bl[K], ba[K]; /* J linear / angular, M^-1 J^T linear / angular */
        T rhs[K], invD[K], rhsP[K];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const bool has = c < nc; /* rows this lane does not have: zero Jacobian, no effect */
            const V3<T> z(T(0), T(0), T(0));
            jl[c] = has ? on[c] : z; ja[c] = has ? orxn[c] : z; bl[c] = jl[c] * invMass; ba[c] = has ? oangN[c] : z;
            jl[NC + c] = has ? ot[c] : z; ja[NC + c] = has ? orxt[c] : z; bl[NC + c] = jl[NC + c] * invMass; ba[NC + c] = has ? oangT[c] : z;
            /* (a row the lane does not have may hold any bit pattern: keep it out of the arithmetic altogether) */
            rhs[c] = c < nc ? oRhsN[c] : T(0); invD[c] = c < nc ? oInvN[c] : T(0); rhsP[c] = c < nc ? oRhsPen[c] : T(0);
            rhs[NC + c] = c < nc ? oRhsT[c] : T(0); invD[NC + c] = c < nc ? oInvT[c] : T(0); rhsP[NC + c] = T(0);
        }
#pragma unroll
        for (int p = 0; p < PMAX; ++p) {
            const T sg = pSign[p];
            const V3<T> z(T(0), T(0), T(0));
            jl[PN + p] = sg != T(0) ? pn[p] * sg : z; ja[PN + p] = sg != T(0) ? prxn[p] * sg : z;
            bl[PN + p] = jl[PN + p] * invMass; ba[PN + p] = sg != T(0) ? pangN[p] * sg : z;
            jl[PT + p] = sg != T(0) ? pt[p] * sg : z; ja[PT + p] = sg != T(0) ? prxt[p] * sg : z;
            bl[PT + p] = jl[PT + p] * invMass; ba[PT + p] = sg != T(0) ? pangT[p] * sg : z;
            const bool mineP = sg != T(0);
            rhs[PN + p] = mineP ? pRhsN[p] : T(0); invD[PN + p] = mineP ? pInvN[p] : T(0); rhsP[PN + p] = pRhsPen[p];
            rhs[PT + p] = mineP ? pRhsT[p] : T(0); invD[PT + p] = mineP ? pInvT[p] : T(0); rhsP[PT + p] = T(0);
        }
        T W[K][K];
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = i; j < K; ++j) {
                const T v = jl[i].dot(bl[j]) + ja[i].dot(ba[j]);
                W[i][j] = v; W[j][i] = v;
            }
        TGB_PAIR_TICK(107);
        /* A row visit, written for the shortest dependent chain (dv -> fma -> max/min -> subtract -> fma into the next row's dv):
           new impulse = clamp(lam + rhs - dv invD) with lam + rhs kept ready from the row's previous visit. */
        T lam[K], dv[K], base[K];
        /* warm starting: the rows start from their scaled impulses of the last step; every row's relative velocity part follows */
#pragma unroll
        for (int c = 0; c < NC; ++c) { lam[c] = oWarmN[c]; lam[NC + c] = oWarmT[c]; }
#pragma unroll
        for (int p = 0; p < PMAX; ++p) { lam[PN + p] = pWarmN[p]; lam[PT + p] = pWarmT[p]; }
#pragma unroll
        for (int i = 0; i < K; ++i) {
            T v = T(0);
#pragma unroll
            for (int j = 0; j < K; ++j) v = fmaf(W[i][j], lam[j], v);
            dv[i] = v; base[i] = lam[i] + rhs[i];
        }
        if (anyPen) {
            /* split impulse: the normal rows alone, their own impulses and relative velocities; a row without it has rhsP = 0 and stays 0 */
            T lamP[K], dvP[K];
#pragma unroll
            for (int i = 0; i < K; ++i) { lamP[i] = T(0); dvP[i] = T(0); }
            for (int it = 0; it < iterations; ++it) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const T nl = rhsP[c] != T(0) ? fmaxf(fmaf(-dvP[c], invD[c], lamP[c] + rhsP[c]), T(0)) : T(0);
                    const T delta = nl - lamP[c];
                    lamP[c] = nl;
#pragma unroll
                    for (int i = 0; i < K; ++i) dvP[i] = fmaf(W[i][c], delta, dvP[i]);
                }
#pragma unroll
                for (int p = 0; p < PMAX; ++p)
                    if (PM == 1 || p < maxP) {
                        const int r_ = PN + p;
                        const T total = dvP[r_] + __shfl_sync(mask, dvP[r_], pSrc[p]);
                        const T nl = rhsP[r_] != T(0) ? fmaxf(fmaf(-total, invD[r_], lamP[r_] + rhsP[r_]), T(0)) : T(0);
                        const T delta = nl - lamP[r_];
                        lamP[r_] = nl;
#pragma unroll
                        for (int i = 0; i < K; ++i) dvP[i] = fmaf(W[i][r_], delta, dvP[i]);
                    }
            }
#pragma unroll
            for (int i = 0; i < K; ++i) { push += bl[i] * lamP[i]; turn += ba[i] * lamP[i]; }
        }
        TGB_PAIR_TICK(108);
        for (int it = 0; it < iterations; ++it) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const T nl = fmaxf(fmaf(-dv[c], invD[c], base[c]), T(0));
                const T delta = nl - lam[c];
                lam[c] = nl; base[c] = nl + rhs[c];
#pragma unroll
                for (int i = 0; i < K; ++i) dv[i] = fmaf(W[i][c], delta, dv[i]);
            }
#pragma unroll
            for (int p = 0; p < PMAX; ++p)
                if (PM == 1 || p < maxP) {
                    const int r_ = PN + p;
                    const T total = dv[r_] + __shfl_sync(mask, dv[r_], pSrc[p]);
                    const T nl = fmaxf(fmaf(-total, invD[r_], base[r_]), T(0));
                    const T delta = nl - lam[r_];
                    lam[r_] = nl; base[r_] = nl + rhs[r_];
#pragma unroll
                    for (int i = 0; i < K; ++i) dv[i] = fmaf(W[i][r_], delta, dv[i]);
                }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int r_ = NC + c;
                const T hi = oMu[c] * lam[c];
                const T nl = fminf(fmaxf(fmaf(-dv[r_], invD[r_], base[r_]), -hi), hi);
                const T delta = lam[c] > T(0) ? nl - lam[r_] : T(0); /* skipped (impulse kept) while the normal impulse is zero */
                lam[r_] += delta; base[r_] = lam[r_] + rhs[r_];
#pragma unroll
                for (int i = 0; i < K; ++i) dv[i] = fmaf(W[i][r_], delta, dv[i]);
            }
#pragma unroll
            for (int p = 0; p < PMAX; ++p)
                if (PM == 1 || p < maxP) {
                    const int r_ = PT + p;
                    const T total = dv[r_] + __shfl_sync(mask, dv[r_], pSrc[p]);
                    const T hi = pMu[p] * lam[PN + p];
                    const T nl = fminf(fmaxf(fmaf(-total, invD[r_], base[r_]), -hi), hi);
                    const T delta = lam[PN + p] > T(0) ? nl - lam[r_] : T(0);
                    lam[r_] += delta; base[r_] = lam[r_] + rhs[r_];
#pragma unroll
                    for (int i = 0; i < K; ++i) dv[i] = fmaf(W[i][r_], delta, dv[i]);
                }
        }
        TGB_PAIR_TICK(109);
#pragma unroll
        for (int i = 0; i < K; ++i) { dLin += bl[i] * lam[i]; dAng += ba[i] * lam[i]; }
        /* the final impulses: own rows back into the rows, pair rows into the persistent manifold (by the lane of body A) */
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (c < nc) { TGB_CTG(c, CT_LAMN) = lam[c]; TGB_CTG(c, CT_LAMT) = lam[NC + c]; }
        if (warmPair != nullptr) {
#pragma unroll
            for (int p = 0; p < PMAX; ++p)
                if (pSign[p] > T(0)) { warmPair[p * TGB_WARM_WORDS] = pKey[p]; warmPair[p * TGB_WARM_WORDS + 1] = lam[PN + p]; warmPair[p * TGB_WARM_WORDS + 2] = lam[PT + p]; }
        }
    } else {
        /* ---- velocity space, the oracle's operations in the oracle's order ---- */
        T oLamN[NC], oLamT[NC], oLamPush[NC];
        T pLamN[PMAX], pLamT[PMAX], pLamPush[PMAX];
        /* warm starting: the rows' scaled impulses of the last step act on the body straight away, row by row in the oracle's order */
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            oLamN[c] = oWarmN[c]; oLamT[c] = oWarmT[c]; oLamPush[c] = T(0);
            if (c < nc) {
                dLin += on[c] * (invMass * oLamN[c]) + ot[c] * (invMass * oLamT[c]);
                dAng += oangN[c] * oLamN[c] + oangT[c] * oLamT[c];
            }
        }
#pragma unroll
        for (int p = 0; p < PMAX; ++p) {
            pLamN[p] = pWarmN[p]; pLamT[p] = pWarmT[p]; pLamPush[p] = T(0);
            if (pSign[p] > T(0)) {
                dLin += pn[p] * (invMass * pLamN[p]) + pt[p] * (invMass * pLamT[p]);
                dAng += pangN[p] * pLamN[p] + pangT[p] * pLamT[p];
            }
            if (pSign[p] < T(0)) {
                dLin -= pn[p] * (invMass * pLamN[p]) + pt[p] * (invMass * pLamT[p]);
                dAng -= pangN[p] * pLamN[p] + pangT[p] * pLamT[p];
            }
        }
        /* one visit of a normal-type row of this lane's own */
        auto ownNormal = [&](int c, T rhsv, T& lamv, V3<T>& lin, V3<T>& ang) {
            T delta = rhsv;
            const T dvv = on[c].dot(lin) + orxn[c].dot(ang);
            delta -= dvv * oInvN[c];
            const T sum = lamv + delta;
            if (sum < T(0)) { delta = -lamv; lamv = T(0); }
            else lamv = sum;
            lin += on[c] * (invMass * delta);
            ang += oangN[c] * delta;
        };
        /* one visit of a pair row's normal constraint; `active` = both lanes of the row work on it (uniform between the two) */
        auto pairNormalReg = [&](int p, T rhsv, T& lamv, V3<T>& lin, V3<T>& ang, bool active) {
            const T dvOwn = pn[p].dot(lin) + prxn[p].dot(ang);
            const T dvOther = __shfl_sync(mask, dvOwn, pSrc[p]);
            const bool isB = pSign[p] < T(0);
            const T dvv = (isB ? dvOther : dvOwn) - (isB ? dvOwn : dvOther);
            T delta = rhsv;
            delta -= dvv * pInvN[p];
            const T sum = lamv + delta;
            T newLam = sum;
            if (sum < T(0)) { delta = -lamv; newLam = T(0); }
            if (active) {
                lamv = newLam;
                if (pSign[p] > T(0)) { lin += pn[p] * (invMass * delta); ang += pangN[p] * delta; }
                if (pSign[p] < T(0)) { lin -= pn[p] * (invMass * delta); ang -= pangN[p] * delta; }
            }
        };
        if (anyPen)
            for (int it = 0; it < iterations; ++it) {
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (c < nc && oRhsPen[c] != T(0)) ownNormal(c, oRhsPen[c], oLamPush[c], push, turn);
#pragma unroll
                for (int p = 0; p < PMAX; ++p)
                    if (PM == 1 || p < maxP) pairNormalReg(p, pRhsPen[p], pLamPush[p], push, turn, pRhsPen[p] != T(0));
            }
        for (int it = 0; it < iterations; ++it) {
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (c < nc) ownNormal(c, oRhsN[c], oLamN[c], dLin, dAng);
#pragma unroll
            for (int p = 0; p < PMAX; ++p)
                if (PM == 1 || p < maxP) pairNormalReg(p, pRhsN[p], pLamN[p], dLin, dAng, pSign[p] != T(0));
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (c < nc && oLamN[c] > T(0)) {
                    const T lo = -oMu[c] * oLamN[c], hi = oMu[c] * oLamN[c];
                    T delta = oRhsT[c];
                    const T dvv = ot[c].dot(dLin) + orxt[c].dot(dAng);
                    delta -= dvv * oInvT[c];
                    const T sum = oLamT[c] + delta;
                    if (sum < lo) { delta = lo - oLamT[c]; oLamT[c] = lo; }
                    else if (sum > hi) { delta = hi - oLamT[c]; oLamT[c] = hi; }
                    else oLamT[c] = sum;
                    dLin += ot[c] * (invMass * delta);
                    dAng += oangT[c] * delta;
                }
#pragma unroll
            for (int p = 0; p < PMAX; ++p)
                if (PM == 1 || p < maxP) {
                    const T dvOwn = pt[p].dot(dLin) + prxt[p].dot(dAng);
                    const T dvOther = __shfl_sync(mask, dvOwn, pSrc[p]);
                    const bool isB = pSign[p] < T(0);
                    const T dvv = (isB ? dvOther : dvOwn) - (isB ? dvOwn : dvOther);
                    const T lamN = pLamN[p];
                    const T lo = -pMu[p] * lamN, hi = pMu[p] * lamN;
                    T delta = pRhsT[p];
                    delta -= dvv * pInvT[p];
                    const T sum = pLamT[p] + delta;
                    T newLam = sum;
                    if (sum < lo) { delta = lo - pLamT[p]; newLam = lo; }
                    else if (sum > hi) { delta = hi - pLamT[p]; newLam = hi; }
                    if (pSign[p] != T(0) && lamN > T(0)) { /* the friction row is skipped (impulse kept) while the normal impulse is zero */
                        pLamT[p] = newLam;
                        if (pSign[p] > T(0)) { dLin += pt[p] * (invMass * delta); dAng += pangT[p] * delta; }
                        if (pSign[p] < T(0)) { dLin -= pt[p] * (invMass * delta); dAng -= pangT[p] * delta; }
                    }
                }
        }
        /* the final impulses: own rows back into the rows, pair rows into the persistent manifold (by the lane of body A) */
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (c < nc) { TGB_CTG(c, CT_LAMN) = oLamN[c]; TGB_CTG(c, CT_LAMT) = oLamT[c]; }
        if (warmPair != nullptr) {
#pragma unroll
            for (int p = 0; p < PMAX; ++p)
                if (pSign[p] > T(0)) { warmPair[p * TGB_WARM_WORDS] = pKey[p]; warmPair[p * TGB_WARM_WORDS + 1] = pLamN[p]; warmPair[p * TGB_WARM_WORDS + 2] = pLamT[p]; }
        }
    }
    ContactSolveOut<T> o;
    o.dLin = dLin; o.dAng = dAng; o.push = push; o.turn = turn;
    return o;
}

template <typename T, bool kFast, bool OBST>
__device__ __forceinline__ ContactSolveOut<T> solveCoupledMem(unsigned mask, unsigned groupMask, T* ct, int rowStride, int nc, bool anyPenOwn, T* pr,
                                                            T* warmPair, int nPairRows, int lane, int groupLane0, const CoupledBody<T>& body,
                                                            int iterations) {
    const V3<T> gn = body.normal();
    const InvInertia<T> Iinv = body.Iinv();
    const T invMass = body.invMass(), muGround = body.mu();
    const int warpLane = groupLane0 + lane;
    const int maxP = __reduce_max_sync(mask, nPairRows);
    const bool anyPen = __any_sync(mask, groupHasPenetration<T>(groupMask, anyPenOwn, pr, nPairRows));
    V3<T> dLin(T(0), T(0), T(0)), dAng(T(0), T(0), T(0)), push(T(0), T(0), T(0)), turn(T(0), T(0), T(0));
    /* warm starting: the rows' initial impulses (as the set-up left them in the rows) act on the bodies straight away */
    for (int c = 0; c < nc; ++c) {
        const T lamN0 = TGB_CTG(c, CT_LAMN), lamT0 = TGB_CTG(c, CT_LAMT);
        const V3<T> n = OBST ? V3<T>(TGB_CTG(c, CT_NRM + 0), TGB_CTG(c, CT_NRM + 1), TGB_CTG(c, CT_NRM + 2)) : gn;
        const V3<T> t(TGB_CTG(c, CT_T + 0), TGB_CTG(c, CT_T + 1), TGB_CTG(c, CT_T + 2));
        const V3<T> angN(TGB_CTG(c, CT_ANGN + 0), TGB_CTG(c, CT_ANGN + 1), TGB_CTG(c, CT_ANGN + 2));
        const V3<T> angT(TGB_CTG(c, CT_ANGT + 0), TGB_CTG(c, CT_ANGT + 1), TGB_CTG(c, CT_ANGT + 2));
        dLin += n * (invMass * lamN0) + t * (invMass * lamT0);
        dAng += angN * lamN0 + angT * lamT0;
    }
    for (int p = 0; p < nPairRows; ++p) {
        const T* row = pr + p * TGB_NPR;
        const int ids = (int)row[PR_IDS];
        const bool isA = lane == (ids & 31), isB = lane == (ids >> 5);
        if (!isA && !isB) continue;
        const V3<T> n(row[PR_N + 0], row[PR_N + 1], row[PR_N + 2]), t(row[PR_T + 0], row[PR_T + 1], row[PR_T + 2]);
        const int po = isB ? PR_RB : PR_RA;
        const V3<T> r(row[po + 0], row[po + 1], row[po + 2]);
        const T lamN0 = row[PR_LAMN], lamT0 = row[PR_LAMT];
        const V3<T> lin = n * (invMass * lamN0) + t * (invMass * lamT0);
        const V3<T> ang = (Iinv * r.cross(n)) * lamN0 + (Iinv * r.cross(t)) * lamT0;
        if (isA) { dLin += lin; dAng += ang; } else { dLin -= lin; dAng -= ang; }
    }
    auto pairVisit = [&](int p, int rhsField, int lamField, bool friction, V3<T>& lin, V3<T>& ang) {
        const bool have = p < nPairRows;
        T* row = pr + (have ? p : 0) * TGB_NPR;
        const int ids = (int)row[PR_IDS];
        const int ia = ids & 31, ib = ids >> 5;
        const bool isA = have && lane == ia, isB = have && lane == ib;
        const int src = (isA || isB) ? groupLane0 + (isA ? ib : ia) : warpLane;
        const int dirField = friction ? PR_T : PR_N;
        const V3<T> d(row[dirField + 0], row[dirField + 1], row[dirField + 2]);
        const int po = isB ? PR_RB : PR_RA;
        const V3<T> r(row[po + 0], row[po + 1], row[po + 2]);
        const V3<T> rxd = r.cross(d);
        const V3<T> angD = Iinv * rxd;
        const T dvOwn = d.dot(lin) + rxd.dot(ang);
        const T dvOther = __shfl_sync(mask, dvOwn, src);
        const T dv = (isB ? dvOther : dvOwn) - (isB ? dvOwn : dvOther);
        const T rhs = row[rhsField];
        const T lamN = row[PR_LAMN];
        const T bound = friction ? row[PR_MU] * lamN : T(0);
        T delta = rhs;
        delta -= dv * row[friction ? PR_INVDIAGT : PR_INVDIAGN];
        const T lam = row[lamField];
        const T sum = lam + delta;
        T newLam = sum;
        if (friction) {
            if (sum < -bound) { delta = -bound - lam; newLam = -bound; }
            else if (sum > bound) { delta = bound - lam; newLam = bound; }
        } else if (sum < T(0)) { delta = -lam; newLam = T(0); }
        const bool skip = friction ? !(lamN > T(0)) : (rhsField == PR_RHSPEN && rhs == T(0));
        __syncwarp(mask); /* both lanes of the row have read its impulse */
        if (isA && !skip) row[lamField] = newLam;
        if (isA && !skip) { lin += d * (invMass * delta); ang += angD * delta; }
        if (isB && !skip) { lin -= d * (invMass * delta); ang -= angD * delta; }
    };
    if (anyPen)
        for (int it = 0; it < iterations; ++it) {
            for (int c = 0; c < nc; ++c) {
                const T rhsPen = TGB_CTG(c, CT_RHSPEN);
                if (rhsPen == T(0)) continue;
                const V3<T> n = OBST ? V3<T>(TGB_CTG(c, CT_NRM + 0), TGB_CTG(c, CT_NRM + 1), TGB_CTG(c, CT_NRM + 2)) : gn;
                const V3<T> rxn(TGB_CTG(c, CT_RXN + 0), TGB_CTG(c, CT_RXN + 1), TGB_CTG(c, CT_RXN + 2));
                const V3<T> angN(TGB_CTG(c, CT_ANGN + 0), TGB_CTG(c, CT_ANGN + 1), TGB_CTG(c, CT_ANGN + 2));
                T delta = rhsPen;
                const T dv = n.dot(push) + rxn.dot(turn);
                delta -= dv * TGB_CTG(c, CT_INVDIAGN);
                const T lam = TGB_CTG(c, CT_LAMPUSH);
                const T sum = lam + delta;
                if (sum < T(0)) { delta = -lam; TGB_CTG(c, CT_LAMPUSH) = T(0); }
                else TGB_CTG(c, CT_LAMPUSH) = sum;
                push += n * (invMass * delta);
                turn += angN * delta;
            }
            __syncwarp(mask);
            for (int p = 0; p < maxP; ++p) pairVisit(p, PR_RHSPEN, PR_LAMPUSH, false, push, turn);
            __syncwarp(mask);
        }
    for (int it = 0; it < iterations; ++it) {
        for (int c = 0; c < nc; ++c) {
            const V3<T> n = OBST ? V3<T>(TGB_CTG(c, CT_NRM + 0), TGB_CTG(c, CT_NRM + 1), TGB_CTG(c, CT_NRM + 2)) : gn;
            const V3<T> rxn(TGB_CTG(c, CT_RXN + 0), TGB_CTG(c, CT_RXN + 1), TGB_CTG(c, CT_RXN + 2));
            const V3<T> angN(TGB_CTG(c, CT_ANGN + 0), TGB_CTG(c, CT_ANGN + 1), TGB_CTG(c, CT_ANGN + 2));
            T delta = TGB_CTG(c, CT_RHSN);
            const T dv = n.dot(dLin) + rxn.dot(dAng);
            delta -= dv * TGB_CTG(c, CT_INVDIAGN);
            const T lam = TGB_CTG(c, CT_LAMN);
            const T sum = lam + delta;
            if (sum < T(0)) { delta = -lam; TGB_CTG(c, CT_LAMN) = T(0); }
            else TGB_CTG(c, CT_LAMN) = sum;
            dLin += n * (invMass * delta);
            dAng += angN * delta;
        }
        __syncwarp(mask);
        for (int p = 0; p < maxP; ++p) pairVisit(p, PR_RHSN, PR_LAMN, false, dLin, dAng);
        __syncwarp(mask);
        for (int c = 0; c < nc; ++c) {
            const T lamN = TGB_CTG(c, CT_LAMN);
            if (!(lamN > T(0))) continue;
            const V3<T> t(TGB_CTG(c, CT_T + 0), TGB_CTG(c, CT_T + 1), TGB_CTG(c, CT_T + 2));
            const V3<T> rxt(TGB_CTG(c, CT_RXT + 0), TGB_CTG(c, CT_RXT + 1), TGB_CTG(c, CT_RXT + 2));
            const V3<T> angT(TGB_CTG(c, CT_ANGT + 0), TGB_CTG(c, CT_ANGT + 1), TGB_CTG(c, CT_ANGT + 2));
            const T mu = OBST ? TGB_CTG(c, CT_MU) : muGround;
            const T lo = -mu * lamN, hi = mu * lamN;
            T delta = TGB_CTG(c, CT_RHST);
            const T dv = t.dot(dLin) + rxt.dot(dAng);
            delta -= dv * TGB_CTG(c, CT_INVDIAGT);
            const T lam = TGB_CTG(c, CT_LAMT);
            const T sum = lam + delta;
            if (sum < lo) { delta = lo - lam; TGB_CTG(c, CT_LAMT) = lo; }
            else if (sum > hi) { delta = hi - lam; TGB_CTG(c, CT_LAMT) = hi; }
            else TGB_CTG(c, CT_LAMT) = sum;
            dLin += t * (invMass * delta);
            dAng += angT * delta;
        }
        __syncwarp(mask);
        for (int p = 0; p < maxP; ++p) pairVisit(p, PR_RHST, PR_LAMT, true, dLin, dAng);
        __syncwarp(mask);
    }
    /* the pair rows' final impulses into the persistent manifold (by the lane of body A) */
    if (warmPair != nullptr)
        for (int p = 0; p < nPairRows; ++p) {
            const T* row = pr + p * TGB_NPR;
            if (lane == ((int)row[PR_IDS] & 31)) {
                warmPair[p * TGB_WARM_WORDS] = row[PR_KEY]; warmPair[p * TGB_WARM_WORDS + 1] = row[PR_LAMN]; warmPair[p * TGB_WARM_WORDS + 2] = row[PR_LAMT];
            }
        }
    ContactSolveOut<T> o;
    o.dLin = dLin; o.dAng = dAng; o.push = push; o.turn = turn;
    return o;
}
#undef TGB_CTG

/* The one call the step kernel makes when a warp has coupled environments (a single call site keeps the hot loop's register
   allocation away from the argument marshalling of four): decides per environment which solver it takes -- by the environment's own
   rows only -- and runs them.  Entered by all 32 lanes of the body warp.  `solved` tells the lane whether its result is in the
   return value (its environment is coupled) or its per-lane solver still has to run. */
template <typename T, bool kFast, bool OBST, int LPE>
__device__ __forceinline__ CoupledOut<T> solveCoupledWarp(T* ct, int rowStride, int nc, bool anyPenOwn, T* pr, T* warmPair, int nPairRows, int lane,
                                                        int groupLane0, const CoupledBody<T>& body, V3<T> vel, V3<T> w, V3<T> extLin, V3<T> extAng, T dt,
                                                        int iterations
#ifdef TGB_STATS
                                                        , unsigned long long* statsP
#endif
                                                        ) {
    const unsigned fullMask = 0xffffffffu;
    const unsigned groupMaskP = (LPE == 32) ? 0xffffffffu : (((1u << LPE) - 1u) << groupLane0);
    const int ncMine = body.mine ? nc : 0;
    const bool penMine = body.mine && anyPenOwn;
    const bool coupled = nPairRows > 0; /* uniform within the environment's group */
    const bool own2 = (__ballot_sync(fullMask, ncMine > 1) & groupMaskP) != 0u;
    const bool own3 = (__ballot_sync(fullMask, ncMine > 2) & groupMaskP) != 0u;
    const bool inReg1 = coupled && !own2 && nPairRows <= 1;
    const bool inReg2 = coupled && !inReg1 && !own3 && nPairRows <= TGB_PAIR_REG_ROWS;
    const unsigned coupledMask = __ballot_sync(fullMask, coupled);
    const unsigned reg1Mask = __ballot_sync(fullMask, inReg1), reg2Mask = __ballot_sync(fullMask, inReg2);
    const unsigned memMask = coupledMask & ~(reg1Mask | reg2Mask);
    __syncwarp(); /* the pair rows and every body's own rows are in shared memory */
    ContactSolveOut<T> o;
    o.dLin = V3<T>(T(0), T(0), T(0)); o.dAng = o.dLin; o.push = o.dLin; o.turn = o.dLin;
    /* the register solvers are entered by the whole warp (full-mask shuffles); lanes of environments that are not theirs come along with
       no rows and drop the result */
    if (reg1Mask != 0u) {
        const ContactSolveOut<T> o1 = solveCoupledReg<T, kFast, OBST, 1, 1>(ct, rowStride, inReg1 ? ncMine : 0, inReg1 && penMine, pr, warmPair,
                                                                            inReg1 ? nPairRows : 0, lane, groupLane0, body, vel, w, extLin, extAng, dt,
                                                                            iterations
#ifdef TGB_STATS
                                                                            , statsP
#endif
                                                                            );
        if (inReg1) o = o1;
    }
    if (reg2Mask != 0u) {
        const ContactSolveOut<T> o2 = solveCoupledReg<T, kFast, OBST, 2, TGB_PAIR_REG_ROWS>(ct, rowStride, inReg2 ? ncMine : 0, inReg2 && penMine, pr,
                                                                                            warmPair, inReg2 ? nPairRows : 0, lane, groupLane0, body, vel,
                                                                                            w, extLin, extAng, dt, iterations
#ifdef TGB_STATS
                                                                                            , statsP
#endif
                                                                                            );
        if (inReg2) o = o2;
    }
    if (coupled && !inReg1 && !inReg2) {
        setupPairRows<T, kFast>(memMask, pr, nPairRows, lane, groupLane0, body, vel, w, extLin, extAng, dt);
        o = solveCoupledMem<T, kFast, OBST>(memMask, groupMaskP, ct, rowStride, ncMine, penMine, pr, warmPair, nPairRows, lane, groupLane0, body, iterations);
    }
    CoupledOut<T> r;
    r.o = o; r.solved = coupled; r.statMem = memMask != 0u;
    return r;
}

} // namespace tgb
