/*
 * tgb_dev.h -- data layout shared by the host side of the C ABI (tgb_capi.cu) and the kernels.
 *
 * HBM layout (one tgb_world, n_env environments, scalar type T = float or double):
 *   state  T[n_env][stride]        env-major records: R bodies x 13 {pos3 (relative to the env frame origin), quat4 xyzw,
 *                                   linvel3 (relative to the env frame velocity), angvel3}, padding to 16 bytes, then C cables x 4
 *                                   {rest length, previous length, preferred length, motor velocity}; stride is padded
 *                                   to a 16-byte multiple so that a group of lanes moves its record with 128-bit
 *                                   accesses and a CTA's records form one contiguous span.
 *   frame  double[n_env][8]         env frame: origin O(3), velocity V(3), controller time, spare (sine controller: 1 when sine_state holds this env's phases).  World position of
 *                                   a body = O + pos, world velocity = V + linvel.  Keeping the common translation of a
 *                                   free-falling structure in double keeps the float state small, so cable stretch
 *                                   (a difference of anchor positions) does not lose digits far from the origin.
 *   ctrl   double[n_env][ctrl_stride] controller parameter rows (tgb_set_controller)
 *   com0   double[n_env][3]         mass-weighted centre of mass at episode start (scores)
 *   status int[n_env]
 *
 * The model (everything identical across environments) is packed once into two small arrays, field-major so that
 * lanes that index consecutive bodies / cables read consecutive shared-memory words:
 *   model_t T[]   : body fields [TGB_NBF][R], shape fields [TGB_NSF][S], cable fields [TGB_NCF][C]
 *   model_i int[] : body ints [TGB_NBI][R], cable ints [TGB_NCI][C], incidence list [n_inc]
 * followed by the per-cable records of the cable phase (TGB_CREC / TGB_CIREC below) at 16-byte aligned offsets.
 */
#ifndef TGB_DEV_H
#define TGB_DEV_H

#include <stdint.h>

/* body T-fields */
enum { BF_INVMASS = 0, BF_GFY, BF_INVIX, BF_INVIY, BF_INVIZ, BF_IX, BF_IY, BF_IZ, BF_MU, BF_RESTITUTION, BF_MASS, BF_FRICTION,
       BF_BOUND /* radius about the body origin that holds every capsule of the body (body-pair culling) */,
       BF_SEGH /* a body that is one cylinder or sphere in the body frame (a plain rod): half length of its core segment; else -1 */,
       BF_SEGR /* and its radius */, TGB_NBF };
/* body int-fields */
enum { BI_SHAPE_BEGIN = 0, BI_SHAPE_COUNT, BI_INC_BEGIN, BI_INC_COUNT, BI_GYRO, TGB_NBI };
/* shape T-fields */
enum { SF_RADIUS = 0, SF_HALFLEN, SF_MARGIN, SF_POSX, SF_POSY, SF_POSZ, SF_ROT0, SF_KIND = SF_ROT0 + 9 /* tgb_shape_kind */, SF_HALFZ /* box */,
       SF_SEGH /* half length of the capsule core segment of the body-pair manifold: max(0, half length - radius), 0 for a sphere */, TGB_NSF };
/* cable T-fields */
enum {
    CF_LAX = 0, CF_LAY, CF_LAZ, CF_LBX, CF_LBY, CF_LBZ, CF_K, CF_C, CF_MAXTENS, CF_TVEL, CF_MINAL, CF_MINRL,
    CF_MRADIUS, CF_MFRICTION, CF_MINERTIA, CF_MAXSTRETCH /* max tension / stiffness */, TGB_NCF
};
/* cable int-fields */
enum { CI_A = 0, CI_B, CI_ACTUATOR, CI_BACKDRIVABLE, TGB_NCI };
/* per-cable constants once more as one record per cable for the step kernel's cable phase (128-bit shared-memory loads): T words
   [la.xyz, k][lb.xyz, damping][tVel, maxTens, maxStretch, minAL][minRL, 0, 0, 0][pad x 4]; the stride of 20 words keeps eight
   consecutive records on distinct banks.  Ints: [body a, body b, actuator | backdrivable << 8, ka | kb << 16] with ka / kb the
   position of this cable's end in the incidence list of body a / b (slot of the impulse exchange, see tgb_step.cuh). */
enum { TGB_CREC = 20, TGB_CIREC = 4 };
/* first word of the cable block in an env record: 16-byte aligned for both scalar types */
#define TGB_CABLE_BASE(R) ((((R) * TGB_NBS) + 3) & ~3)
/* per-cable state words in the env record */
enum { CS_REST = 0, CS_PREV, CS_PREF, CS_MOTORVEL, TGB_NCS };
/* per-body state words in the env record */
enum { TGB_NBS = 13 };
/* per-contact solver row words kept in shared memory */
enum {
    CT_RXN = 0, CT_ANGN = 3, CT_INVDIAGN = 6, CT_RHSN, CT_RHSPEN, CT_LAMN, CT_LAMPUSH, CT_T = 11, CT_RXT = 14, CT_ANGT = 17,
    CT_INVDIAGT = 20, CT_RHST, CT_LAMT,
    CT_KEY /* identity of the manifold point across steps (persistent manifold, warm starting): see TGB_WARM_* */,
    TGB_NCT = 25 /* odd row stride: rows of neighbouring lanes on distinct banks */,
    /* worlds with static obstacles: the row's own normal and friction coefficient behind the common part */
    CT_NRM = CT_KEY + 1, CT_MU = CT_NRM + 3, TGB_NCT_OBST = 29
};
/* Persistent manifold: per environment and dynamic body the (key, normal impulse, friction impulse) of every row of the last world
   step, [cap_contacts][row bodies][3], then [pair_cap][3] for the rows between two bodies; key < 0 = empty.  A point keeps its
   identity through its slot: ground: (shape of the body, rim point / corner); obstacle: TGB_WARM_OBST + (obstacle, shape, sample);
   pair: (body A + 32 body B, shape of A, shape of B, point).  In shared memory during a launch, in HBM (tgb_step_args.warm) between
   launches, so that results do not depend on how an episode is cut into launches. */
enum { TGB_WARM_WORDS = 3, TGB_WARM_OBST = 4096 };
#define TGB_WARM_STATE_WORDS(cap, rowBodies, pairCap) (TGB_WARM_WORDS * ((cap) * (rowBodies) + (pairCap)))
/* Row of a contact between two bodies (tgb_pairs.cuh), T words in shared memory.  The two sides' Jacobian blocks are not stored: each
   side's lane re-forms r x n, Iinv (r x n) from its own r and the inverse inertia in its registers.  Between detection and row
   set-up PR_RA / PR_RB hold the contact points on the two surfaces (frame-relative), PR_LAMN the distance, PR_LAMPUSH the
   combined restitution, PR_INVDIAGT / PR_RHST the point's warm-start impulses (normal, friction; already scaled). */
enum {
    PR_N = 0, PR_RA = 3, PR_RB = 6, PR_T = 9, PR_INVDIAGN = 12, PR_RHSN, PR_RHSPEN, PR_LAMN, PR_LAMPUSH, PR_INVDIAGT, PR_RHST, PR_LAMT,
    PR_MU, PR_IDS /* body A + 32 * body B */, PR_KEY /* identity of the point across steps (warm starting) */, TGB_NPR
};
/* frame doubles */
enum { FR_O = 0, FR_V = 3, FR_CTRLTIME = 6, FR_SPARE = 7, TGB_NFR = 8 };

#define TGB_STATUS_NONFINITE 1
#define TGB_STATUS_HEIGHT 2
#define TGB_STATUS_CPG_INEFFICIENT 4
/* doubles of CPG controller state per environment */
#define TGB_CPG_STATE_DOUBLES(C) (6 * (C) + 2)

struct tgb_step_args {
    void* state;          /* T[n_env][stride] */
    double* frame;        /* [n_env][TGB_NFR] */
    const double* ctrl;   /* [n_env][ctrl_stride] or NULL */
    double* sine_state;   /* sine controller, single precision: [n_env][24][2] (sin, cos) of each cable's phase; valid when the frame's
                             spare word is non-zero */
    int* status;          /* [n_env] */
    const void* model_t;  /* packed T fields */
    const int* model_i;   /* packed int fields */
    void* max_tension;    /* T[n_env][C] running maximum of the logged tension (tgb_track_max_tension), or NULL */
    unsigned long long* stats; /* instrumentation variant only (-DTGB_STATS, tools/), else unused */
    int n_env, n_steps;   /* a launch covers environments [env_begin, n_env) */
    int env_begin;
    int phases;           /* bit0: world half of tgSimulation::step, bit1: models half (controllers, motors, cable forces) */
    int wide;             /* 1: the four-warp layout (body warp + shared cable phase, tgb_step.cuh); CTAs of 128 threads hold 32 / lanes environments */
    int R, C, S, n_inc;
    int max_inc;          /* longest incidence list of a body */
    int stride;           /* T words per env record */
    int cap_contacts;     /* contact slots per body */
    int frame_body;       /* body whose position/velocity the env frame follows */
    int ctrl_kind, ctrl_stride;
    int iterations;       /* PGS iterations */
    int ground_kind;
    int model_t_words, model_i_words;
    int model_crec_off, model_cirec_off; /* word offsets of the per-cable records inside model_t / model_i */
    int n_dyn;                           /* bodies with mass: worlds with obstacles keep contact rows for these only */
    int n_obst, model_obst_off;          /* static obstacles: n_obst pairs (static body, its box shape) at this word offset of model_i */
    int n_pairs, model_pair_off;         /* pairs of bodies that can collide (at least one dynamic): body A + 32 * body B, A < B, at this word offset of model_i */
    int pair_cap;                        /* rows of body-pair contacts per environment; 0 = bodies do not collide with each other */
    void* warm;                          /* T[n_env][TGB_WARM_STATE_WORDS]: the persistent manifolds between launches, or NULL */
    double warm_factor;                  /* btContactSolverInfo::m_warmstartingFactor (0.85); 0 = every solve starts from zero */
    float* min_clearance;                /* [n_env] running minimum of the surface distance between two bodies' capsules, or NULL */
    /* Launches over a list of environments (tgb_capi.cu, "tail split"): slot i of the grid simulates environment env_list[i] for
       i < min(*env_list_count, env_list_slots); CTAs beyond the count leave at once.  NULL: slot i is environment env_begin + i. */
    const int* env_list;
    const int* env_list_count;
    int env_list_slots;
    int* coupled_steps;                  /* [n_env] steps of this launch in which the environment had rows between two bodies or a body
                                            with three or more contact points (the steps that hold a warp up) in the low 20 bits, those
                                            the number of such steps in a row up to the launch's last step (capped at 2047) above them; or NULL */
    double dt;
    double gn[3], goff;   /* ground surface { x : gn.x = goff } */
    double ground_restitution;
    double height_lo, height_hi; /* status bit1 when the mass centre leaves [lo, hi]; lo > hi disables */

    /* CPG + impedance controller (TGB_CTRL_CPG_IMPEDANCE); ctrl rows are [freq, radius, n_pairs x (weight, phase)] */
    int cpg_n_couplings, cpg_n_pairs, cpg_max_steps, cpg_score_body;
    const int* cpg_topo;      /* begin[C + 1], target[n_couplings], pair[n_couplings] */
    const double* cpg_imp;    /* [C][4] offset tension, length stiffness, velocity stiffness, control length */
    double* cpg_state;        /* [n_env][6 C + 2]: (phase, radius, radius rate) per node, update time, cable control time, then per
                                 cable the tgPIDController state (previous error, integrated error) and the last commanded tension */
    int cpg_cable_control;    /* tgb_cpg_desc.cable_control */
    double cpg_pid[3];        /* kP, kI, kD */
    double cpg_r_const, cpg_d_min, cpg_d_max, cpg_desc, cpg_control_time, cpg_cable_step, cpg_height_lo, cpg_height_hi;
};

#endif
