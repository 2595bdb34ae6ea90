/*
 * tg_oracle.h -- CPU oracle for the batched tensegrity step (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * A plain double-precision, one-environment-at-a-time restatement of the reference's
 * per-step path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use anything in oracle/.  The shipped library
 * (ntrtsim_b200/csrc) never includes, links or calls it.
 *
 * What it follows (paths relative to the NTRTsim tree):
 *   step order            src/core/tgSimulation.cpp:193-225   (world first, then models)
 *   cable force           src/core/tgBulletSpringCable.cpp:90-141
 *   anchors               src/core/tgBulletSpringCableAnchor.cpp:76-87
 *   tension / length      src/core/tgBulletSpringCable.cpp:143-155
 *   motors                src/core/tgBasicActuator.cpp:137-225, src/core/tgKinematicActuator.cpp:135-225
 *   controllers           src/controllers/tgTensionController.cpp:50-88,
 *                         src/examples/craterEscape/EscapeController.cpp:100-117,245-291,
 *                         src/examples/motorModel/tsTestRig.cpp:170-199,
 *                         src/controllers/tgImpedanceController.cpp:98-185, src/controllers/tgPIDController.cpp:80-109,
 *                         src/examples/learningSpines/{BaseSpineCPGControl,tgCPGActuatorControl,tgCPGCableControl}.cpp,
 *                         src/util/{CPGEquations,CPGNode}.cpp (Boost.Odeint integrate() restated in tg_odeint.h),
 *                         src/examples/IROS_2015/TetraSpineStatic/SerializedSpineControl.cpp:174-205,262-371
 *   contact cables        src/core/tgBulletContactSpringCable.cpp:97-190 (step, calculateAndApplyForce, updateManifolds, updateAnchorList,
 *                         pruneAnchors, updateAnchorPositions), src/core/tgBulletSpringCableAnchor.cpp:89-334 (sliding anchors) -- the cable
 *                         logic restated as written and PINNED against those two sources compiled in place (oracle/ref/cc_hybrid.py,
 *                         tests/golden/contact_cable_demo.npz: 1e-9 over 10000 steps, identical anchor lists); Bullet's ghost object
 *                         and manifolds under it restated geometrically (include/ntrtsim_b200.h "Contact cables"): unpinned against
 *                         Bullet; the whole also pinned by the reference suite's MuscleNP_test (momentum, energy) and by mechanics
 *   world step            Bullet 2.82 btDiscreteDynamicsWorld::stepSimulation(dt,1,dt) as reached from
 *                         src/core/tgWorldBulletPhysicsImpl.cpp:209-221 -- Bullet itself is NOT in the
 *                         reference tree (conf/default/bullet.conf.default:34,44); its semantics are
 *                         restated from SURVEY.md Appendix B.
 *
 * PARITY STATUS: pieces (1) cables and (2) motors/controllers are pinned against the reference's
 * own sources compiled in place (oracle/_ref, see oracle/ref/README.md) through tests/golden/.
 * Piece (3) rigid-body integration is pinned by analytic invariants; piece (4) contact has no per-step
 * pin ("parity unpinned" step by step): Bullet's GJK/persistent-manifold narrowphase and its Dantzig LCP
 * solver are replaced here by an analytic cylinder-rim / sphere / box-corner manifold against the ground, a
 * sample-point manifold against the box shapes of static bodies (obstacles), and fixed-iteration projected
 * Gauss-Seidel; its friction and normal rows are checked against mechanics (sliding ball, brick on an incline:
 * tests/test_cpu.py).  The whole path, contact included, is pinned at the level of a percent by the reference
 * suite's own known-answer test test_integration/ICRA2015Tests/ICRA2015_test.cpp (maxima of six cable
 * tensions over 60000 steps against hardware, with the reference simulator's own errors as bounds):
 * tests/test_golden.py::test_icra2015_hardware_comparison_of_the_reference_test_suite.
 */
#ifndef TG_ORACLE_H
#define TG_ORACLE_H

#include "../include/ntrtsim_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tgo_world tgo_world;

/* One environment. Same descriptors as the device library so both sides see identical inputs. */
tgo_world* tgo_create(const tgb_world_cfg* cfg, const tgb_ground_desc* ground);
void       tgo_destroy(tgo_world* w);
int        tgo_add_model_flat(tgo_world* w, const tgb_body_desc* bodies, int n_bodies,
                              const tgb_shape_desc* shapes, int n_shapes,
                              const tgb_cable_desc* cables, int n_cables);
/* contact cables (tgBulletContactSpringCable, src/core/tgBulletContactSpringCable.cpp:97-190; see include/ntrtsim_b200.h): body
   indices relative to the model added last.  obs [n][5]; anchors [TGB_CCABLE_MAX_SLIDING][7] */
int  tgo_add_contact_cables(tgo_world* w, const tgb_contact_cable_desc* cables, int n);
int  tgo_num_contact_cables(const tgo_world* w);
int  tgo_contact_cable_points(const tgo_world* w, int cable, const double* chain, int n, double* out); /* out [<= 64][8] */
int  tgo_set_contact_cable_control_input(tgo_world* w, const double* input, double dt); /* [n_contact_cables] */
void tgo_set_batch_contact_cables(const tgb_contact_cable_desc* cables, int n); /* for the worlds of tgo_rollout_batch; 0 clears */
void tgo_get_contact_cable_obs(const tgo_world* w, double* obs5);
void tgo_get_contact_cable_anchors(const tgo_world* w, int cable, double* out);
int  tgo_num_bodies(const tgo_world* w);
int  tgo_num_cables(const tgo_world* w);

void tgo_set_rod_state(tgo_world* w, const double* state13);   /* [n_bodies][13] */
void tgo_get_rod_state(const tgo_world* w, double* state13);
void tgo_set_cable_state(tgo_world* w, const double* rest_prev); /* [n_cables][2] */
void tgo_get_cable_obs(const tgo_world* w, double* obs4);        /* [n_cables][4] */

int  tgo_set_controller(tgo_world* w, int kind, const double* params, int n_params);
void tgo_track_max_tension(tgo_world* w, int enable);
void tgo_get_max_tension(const tgo_world* w, double* out);
int  tgo_set_control_input(tgo_world* w, const double* input, double dt); /* [n_cables] */
/* CPG + impedance controller of the learningSpines examples; params = one row as in tgb_set_cpg_controller */
int  tgo_set_cpg_controller(tgo_world* w, const tgb_cpg_desc* desc, const double* params, int n_params);
void tgo_cpg_update(tgo_world* w, double dt);          /* one CPGEquations::update on its own */
int  tgo_get_status(const tgo_world* w);              /* bit1 height window left (bogus), bit2 inefficient CPG */
void tgo_get_cpg_state(const tgo_world* w, double* x); /* [n_nodes][3] phase, radius, radius rate */

/* n_steps x tgSimulation::step(dt). Returns 0, or TGB_ERR_INVALID_ARGUMENT when dt <= 0. */
int  tgo_step(tgo_world* w, double dt, int n_steps);

/* The two halves of one step, exposed for the reference-source harness (oracle/ref): */
void tgo_world_step(tgo_world* w, double dt);     /* Bullet stepSimulation(dt,1,dt) restatement */
void tgo_models_step(tgo_world* w, double dt);    /* controllers + motors + cable impulses */

double tgo_get_time(const tgo_world* w);
void   tgo_mark_episode_start(tgo_world* w);
void   tgo_get_scores(const tgo_world* w, double* scores2);
int    tgo_num_contacts(const tgo_world* w);      /* contacts generated in the last world step */
int    tgo_num_pair_contacts(const tgo_world* w); /* of which rows between two bodies */
/* rows of body-pair contacts kept per environment (same meaning as tgb_set_pair_contact_cap; default 6, 0 = bodies do not collide) */
void   tgo_set_pair_contact_cap(tgo_world* w, int cap);
void   tgo_set_default_pair_contact_cap(int cap);
/* warm starting factor of the contact solver (btContactSolverInfo::m_warmstartingFactor, 0.85): every manifold point starts a solve
   from its impulses of the previous step times this; 0 = every solve starts from zero (no persistent manifold). */
void   tgo_set_warm_start(tgo_world* w, double factor);  /* for worlds created afterwards (tgo_rollout_batch creates its own) */
/* smallest surface distance between the collision shapes of two different bodies seen by any world step so far */
double tgo_get_min_clearance(const tgo_world* w);
void   tgo_reset_min_clearance(tgo_world* w);

/* Batched convenience for tests and the CPU baseline: run `n_env` independent environments that share one
   model, each from its own rod state / cable state / controller row, for n_steps; OpenMP-free, `n_threads`
   std::threads.  Outputs as in tgb_get_rod_state / tgb_get_cable_obs. */
int tgo_rollout_batch(const tgb_world_cfg* cfg, const tgb_ground_desc* ground,
                      const tgb_body_desc* bodies, int n_bodies,
                      const tgb_shape_desc* shapes, int n_shapes,
                      const tgb_cable_desc* cables, int n_cables,
                      int ctrl_kind, const double* ctrl_params, int params_per_env,
                      const double* setup_input, double setup_dt,
                      int n_env, const double* rod_state_in, const double* cable_state_in,
                      double dt, int n_steps, int n_threads,
                      double* rod_state_out, double* cable_obs_out,
                      const tgb_cpg_desc* cpg_desc, double* scores_out, int* status_out);

#ifdef __cplusplus
}
#endif
#endif
