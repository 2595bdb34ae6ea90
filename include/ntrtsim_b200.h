/*
 * ntrtsim_b200.h -- C ABI of the B200-native batched tensegrity step.
 *
 * One `tgb_world` holds N independent copies ("environments") of one flattened
 * NTRTsim world: rigid bodies (rods / compounds of cylinders), two-point
 * spring-cables with their actuators, one ground, one gravity value.  Every
 * call below replaces a piece of the per-step path of the reference
 * (paths relative to the NTRTsim source tree):
 *
 *   tgb_create / tgb_destroy     src/core/tgWorld.cpp:75-96  (tgWorld ctor/reset -> new tgWorldBulletPhysicsImpl)
 *                                 src/core/tgWorldBulletPhysicsImpl.cpp:114-152 (gravity, ground body)
 *   tgb_add_model_flat           src/tgcreator/tgStructureInfo.cpp:277-289 (buildInto: bodies + cables reach the world)
 *                                 src/core/tgBulletUtil.cpp:39-77 (createRigidBody), src/core/tgSpringCable.cpp:33-60
 *   tgb_finalize                 src/core/tgSimulation.cpp:65-82 (addModel -> setup): state becomes steppable
 *   tgb_step                     src/core/tgSimulation.cpp:193-225 (world step, then model/actuator/cable step)
 *                                 = tgWorldImpl::step (src/core/tgWorldImpl.h:56) + tgBasicActuator::step
 *                                 (src/core/tgBasicActuator.cpp:96-113) for every environment, n_steps times
 *   tgb_set_control_input        src/core/tgBasicActuator.cpp:149-168 (setControlInput(input, dt) -> moveMotors)
 *   tgb_set_controller           device-side forms of the shipped observers:
 *                                 src/examples/craterEscape/EscapeController.cpp:100-117,245-291 (sine)
 *                                 src/examples/SUPERball/controllers/T6TensionController.cpp:69-83 (tension)
 *                                 src/examples/motorModel/tsTestRig.cpp:181-188 (set-point / motor torque)
 *   tgb_get_rod_state            src/core/tgBaseRigid.cpp:74-103 (centerOfMass / orientation)
 *   tgb_get_cable_obs            src/core/tgBulletSpringCable.cpp:143-155 (getActualLength / getTension),
 *                                 src/core/tgSpringCable.cpp:65-68 (getRestLength)
 *   tgb_get_scores               src/examples/craterEscape/EscapeController.cpp:119-135 (displacement score)
 *
 * Plain pointers and sizes only; all host-side numbers are double (the
 * reference is double precision throughout); the device computes in the
 * precision chosen in tgb_world_cfg.  No call falls back to a CPU
 * implementation: without a CUDA device tgb_create fails with TGB_ERR_CUDA.
 *
 * Thread-safety: like the reference (single-threaded, no re-entrancy) a
 * tgb_world must be used from one host thread at a time.
 */
#ifndef NTRTSIM_B200_H
#define NTRTSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGB_ABI_VERSION 6

typedef enum {
    TGB_OK = 0,
    TGB_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument in the reference (dt<=0, negative input, bad config) */
    TGB_ERR_STATE = 2,            /* call order (e.g. step before finalize) */
    TGB_ERR_CUDA = 3,             /* no device / CUDA runtime failure */
    TGB_ERR_UNSUPPORTED = 4,      /* model outside the compiled limits */
    TGB_ERR_RUNTIME = 5           /* std::runtime_error in the reference (controllers) */
} tgb_status;

typedef enum { TGB_F32 = 0, TGB_F64 = 1 } tgb_precision;

typedef enum {
    TGB_GROUND_EMPTY = 0, /* tgEmptyGround: no ground body (tgWorldBulletPhysicsImpl.cpp:125-128) */
    TGB_GROUND_BOX = 1,   /* tgBoxGround: top face of the static box, friction 0.5 / restitution 0 (Bullet defaults) */
    TGB_GROUND_PLANE = 2  /* tgPlaneGround: btStaticPlaneShape, friction/restitution from its Config */
} tgb_ground_kind;

typedef struct {
    double gravity;        /* tgWorld::Config::gravity; acts along -Y */
    int32_t n_env;         /* number of independent environments held by this world */
    int32_t precision;     /* tgb_precision: arithmetic type of the device state */
    int32_t device;        /* CUDA device ordinal */
    int32_t solver_iterations; /* contact PGS iterations, 0 -> 10 (Bullet default numIterations) */
} tgb_world_cfg;

typedef struct {
    int32_t kind;          /* tgb_ground_kind */
    double normal[3];      /* unit normal of the contact surface (world) */
    double offset;         /* surface is { x : normal.x = offset } */
    double friction;
    double rolling_friction;
    double restitution;
    double margin;         /* collision margin of the ground shape (0.04 for the box) */
} tgb_ground_desc;

/* One rigid body = btRigidBody of the reference (a rod, or an auto-compound of rods). */
typedef struct {
    double mass;           /* already squeezed through float, tgBulletUtil.cpp:39-42 ; 0 => static */
    double inertia[3];     /* local diagonal inertia from calculateLocalInertia */
    double pos[3];         /* initial centre-of-mass position */
    double quat[4];        /* initial orientation, (x,y,z,w) */
    double friction;
    double rolling_friction;
    double restitution;
    int32_t shape_begin;   /* first child collision shape in the shape array */
    int32_t shape_count;
    int32_t gyroscopic;    /* BT_ENABLE_GYROPSCOPIC_FORCE, tgRigidInfo.cpp:118 */
    int32_t reserved;
} tgb_body_desc;

typedef enum {
    TGB_SHAPE_CYLINDER = 0, /* btCylinderShape: flat-ended cylinder along its local Y (tgRodInfo.cpp:105-106) */
    TGB_SHAPE_SPHERE = 1,   /* btSphereShape (tgSphereInfo.cpp:100-115): a point with collision margin = radius; set
                               half_length = margin = radius */
    TGB_SHAPE_BOX = 2       /* btBoxShape(width, length / 2, height) (tgBoxInfo.cpp:100-118): half extents radius (local X),
                               half_length (local Y, the pair's axis) and half_extent_z.  A box shape of a STATIC body (mass 0: the
                               obstacle models of src/models/obstacles are tgBoxInfo structures with density 0) is an obstacle:
                               every dynamic body of the environment collides with it, with the friction and restitution of the
                               two bodies multiplied.  Static cylinders / spheres only anchor cables (tsTestRig), nothing collides
                               with them. */
} tgb_shape_kind;

/* One collision primitive of a body. */
typedef struct {
    double radius;
    double half_length;
    double margin;         /* cylinder, box: Bullet setSafeMargin, min(0.04, 0.1 * smallest half extent); sphere: the radius */
    double pos[3];         /* child transform in the body frame */
    double quat[4];
    int32_t kind;          /* tgb_shape_kind */
    int32_t reserved;
    double half_extent_z;  /* box only (ABI version 3) */
} tgb_shape_desc;

typedef enum {
    TGB_ACT_PASSIVE = 0,   /* tgBulletSpringCable with no motor */
    TGB_ACT_BASIC = 1,     /* tgBasicActuator */
    TGB_ACT_KINEMATIC = 2  /* tgKinematicActuator */
} tgb_actuator_kind;

/* One two-anchor spring-cable plus its actuator (tgBulletSpringCable + tgSpringCableActuator::Config). */
typedef struct {
    int32_t body_a, body_b;
    double local_a[3], local_b[3]; /* anchor points in the body frames (tgBulletSpringCableAnchor.cpp:54-55) */
    double stiffness, damping;
    double rest_length;    /* initial: |p2-p1| - pretension/k (tgSpringCable.cpp:49-59) */
    double prev_length;    /* initial: == rest_length (first-step damping artefact) */
    int32_t actuator;      /* tgb_actuator_kind */
    int32_t backdrivable;
    double max_tension, target_velocity, min_actual_length, min_rest_length;
    double motor_radius, motor_friction, motor_inertia; /* tgKinematicActuator::Config */
} tgb_cable_desc;

typedef enum {
    TGB_CTRL_NONE = 0,
    /* params per env: [C] preferred lengths; every step setControlInput(pref, dt) (tsTestRig.cpp:181-188) */
    TGB_CTRL_SETPOINT = 1,
    /* params per env: [1] tension set point; tgTensionController::control(dt, T) on every cable
       (T6TensionController.cpp:69-83 -> tgTensionController.cpp:50-88) */
    TGB_CTRL_TENSION = 2,
    /* params per env: [8][4] = amplitude, angular frequency, phase change, dc offset per cluster, already
       passed through transformActions; followed by [1] initial length L0.  EscapeController.cpp:100-117,245-291 */
    TGB_CTRL_ESCAPE_SINE = 3,
    /* params per env: [C] desired motor torque, re-applied every step (tsTestRig.cpp:181-188, kinematic) */
    TGB_CTRL_TORQUE = 4,
    /* CPG + impedance control of the learningSpines examples, installed with tgb_set_cpg_controller:
       BaseSpineCPGControl::onStep (BaseSpineCPGControl.cpp:253-279) -> CPGEquations::update (util/CPGEquations.cpp:231-267),
       tgCPGActuatorControl::onStep (tgCPGActuatorControl.cpp:79-100) -> tgImpedanceController::control
       (controllers/tgImpedanceController.cpp:151-185) -> tgTensionController::control (tgTensionController.cpp:90-114);
       or, with tgb_cpg_desc.cable_control = 1, tgCPGCableControl::onStep -> tgPIDController -> tgKinematicActuator */
    TGB_CTRL_CPG_IMPEDANCE = 5,
    /* SerializedSpineControl of the ICRA 2015 hardware comparison (src/examples/IROS_2015/TetraSpineStatic/SerializedSpineControl.cpp:
       174-205 applyImpedanceControlGeneric, :262-371 onStep): every step, every tgBasicActuator cable runs
       tgImpedanceController::controlTension(actuator, dt, control length, offset tension, target velocity) with the target velocity
       offset_speed + cpg_amplitude * sin(simTime * cpg_frequency + phase), simTime advancing in jumps of 1 / update_frequency; then a
       load is applied to up to three bodies at markers (btRigidBody::applyForce, consumed by the next world step): a force vector
       scaled by a ramp (0 before t0, linear up to t1, constant to t2, linear down to t3).  params per env, TGB_SPINE_HEADER + 5 C:
         [0..3]  offset_speed, cpg_amplitude, cpg_frequency, update_frequency
         [4..7]  t0, t1, t2, t3 of the load ramp       [8..10] load vector at full scale      [11] number of loaded bodies (<= 3)
         [12..26] per loaded body: body index, marker offset in the body frame (3), factor on the load vector
         then per cable: offset tension, length stiffness, velocity stiffness, control length, phase */
    TGB_CTRL_SERIALIZED_SPINE = 6
} tgb_controller_kind;
#define TGB_SPINE_HEADER 27

/* Static part of a CPG + impedance controller: one CPG node per cable (node i drives cable i, the order of
   tgCPGActuatorControl::assignNodeNumber), the coupling lists of tgCPGActuatorControl::setConnectivity in CSR form, and the
   impedance gains of each cable.  The per-environment parameter row is
       [0] frequency parameter (CPGNode frequencyOffset = frequencyScale), [1] radius parameter (radiusOffset = radiusScale),
       then n_pairs x (coupling weight, coupling phase offset);  coupling_pair[] selects the pair of each coupling. */
typedef struct {
    int32_t n_nodes;
    int32_t n_couplings;
    const int32_t* coupling_begin;  /* [n_nodes + 1] */
    const int32_t* coupling_target; /* [n_couplings] node index */
    const int32_t* coupling_pair;   /* [n_couplings] index into the env's (weight, phase) pairs */
    int32_t n_pairs;
    int32_t max_steps;              /* CPGEquations(maxSteps): more right-hand-side evaluations in one update set status bit2 */
    const double* impedance;        /* [n_nodes][4]: offset tension, length stiffness, velocity stiffness, control length */
    double r_const, d_min, d_max;   /* CPGNode params[4..6]: 20, 0, 5 (tgCPGActuatorControl.cpp:102-119) */
    double descending_command;      /* 2.0 (BaseSpineCPGControl.cpp:261) */
    double control_time;            /* BaseSpineCPGControl::Config::controlTime */
    double cable_control_step;      /* tgCPGActuatorControl::m_controlStep, 1e-4 */
    int32_t score_body;             /* body whose centre of mass is scored and height-checked (the segment's compound body) */
    /* Per-cable controller attached by setupCPGs:
         0  tgCPGActuatorControl (BaseSpineCPGControl.cpp:211-213; tgCPGActuatorControl.cpp:79-100): tgBasicActuator cables only; the
            impedance controller is called with the accumulated control time as dt
         1  tgCPGCableControl (KinematicSpineCPGControl.cpp:69-76; tgCPGCableControl.cpp:63-98): tgKinematicActuator cables run
            tgImpedanceController::control(tgBasicController&, ...) (tgImpedanceController.cpp:105-150) through a tgPIDController
            (controllers/tgPIDController.cpp:80-98) whose output is the desired motor torque, and between control steps the PID
            keeps tracking the last commanded tension; tgBasicActuator cables are controlled as in 0 but with the step's dt */
    int32_t cable_control;
    double height_lo, height_hi;    /* 1.0 and 25 (BaseSpineCPGControl.cpp:271-278): outside -> status bit1, score -1 */
    double pid_kp, pid_ki, pid_kd;  /* cable_control 1: signed gains of tgPIDController::Config (tension control: -p, -i, -d) */
} tgb_cpg_desc;

#define TGB_ESCAPE_PARAMS 33

/* Per-rod state record exchanged with the host: pos(3) quat(4, xyzw) linvel(3) angvel(3). */
#define TGB_ROD_STATE_DOUBLES 13
/* Per-cable observation record: rest length, actual length, tension (no damping term), velocity (tgKinematicActuator: motor
   velocity m_motorVel; otherwise the cable's m_velocity = (length - previous length) / dt of the last force evaluation). */
#define TGB_CABLE_OBS_DOUBLES 4

typedef struct tgb_world tgb_world;

int          tgb_abi_version(void);
const char*  tgb_last_error(void);
int          tgb_device_count(void);

/* tgSpringCableActuator::Config::hist = true for every cable (tgSpringCableActuator.h:104-108, tgBasicActuator.cpp:227-241): keep the
   running maximum of the tension logged after every step -- the quantity BaseSpineModelLearning::getStringMaxTensions reads from the
   tension history (ICRA2015_test.cpp:119).  Enabling (or tgb_mark_episode_start) clears it.  out: [env_count][n_cables] doubles. */
tgb_status tgb_track_max_tension(tgb_world* w, int32_t enable);
tgb_status tgb_get_max_tension(tgb_world* w, int32_t env_begin, int32_t env_count, double* out);

tgb_status tgb_create(const tgb_world_cfg* cfg, const tgb_ground_desc* ground, tgb_world** out);
tgb_status tgb_destroy(tgb_world* w);

/* Append one flattened model (indices in `cables` are relative to `bodies`). May be called several times
   before tgb_finalize, like tgSimulation::addModel. */
tgb_status tgb_add_model_flat(tgb_world* w,
                              const tgb_body_desc* bodies, int32_t n_bodies,
                              const tgb_shape_desc* shapes, int32_t n_shapes,
                              const tgb_cable_desc* cables, int32_t n_cables);

/* Allocate the device buffers and replicate the model's initial state into all n_env environments. */
tgb_status tgb_finalize(tgb_world* w);

tgb_status tgb_num_bodies(const tgb_world* w, int32_t* out);
tgb_status tgb_num_cables(const tgb_world* w, int32_t* out);
tgb_status tgb_num_envs(const tgb_world* w, int32_t* out);

/* state: [env_count][n_bodies][13] doubles (host). */
tgb_status tgb_set_rod_state(tgb_world* w, int32_t env_begin, int32_t env_count, const double* state);
tgb_status tgb_get_rod_state(tgb_world* w, int32_t env_begin, int32_t env_count, double* state);

/* rest_prev: [env_count][n_cables][2] doubles (rest length, previous length). */
tgb_status tgb_set_cable_state(tgb_world* w, int32_t env_begin, int32_t env_count, const double* rest_prev);
/* obs: [env_count][n_cables][4] doubles (rest, length, tension, motor velocity). */
tgb_status tgb_get_cable_obs(tgb_world* w, int32_t env_begin, int32_t env_count, double* obs);

/* Install a device-side controller. params: [env_count][params_per_env] doubles; if broadcast != 0 `params`
   holds ONE row that is copied to every environment in the range. */
tgb_status tgb_set_controller(tgb_world* w, int32_t kind, const double* params, int32_t params_per_env,
                              int32_t env_begin, int32_t env_count, int32_t broadcast);

/* Install the CPG + impedance controller (TGB_CTRL_CPG_IMPEDANCE). params_per_env must be 2 + 2 * n_pairs. The CPG state
   (phase, radius, radius rate of every node) of the environments in the range restarts at zero, like a new CPGEquations. */
tgb_status tgb_set_cpg_controller(tgb_world* w, const tgb_cpg_desc* desc, const double* params, int32_t params_per_env,
                                  int32_t env_begin, int32_t env_count, int32_t broadcast);

/* CPG state of the environments in the range: [env_count][n_nodes][3] doubles (phase, radius, radius rate), the XVars of
   CPGEquations::getXVars (util/CPGEquations.cpp:118-131). */
tgb_status tgb_get_cpg_state(tgb_world* w, int32_t env_begin, int32_t env_count, double* out);

/* tgBasicActuator::setControlInput(input, dt) on every cable of every env in the range, once, now.
   input: [n_cables] (broadcast != 0) or [env_count][n_cables]. Negative input -> TGB_ERR_INVALID_ARGUMENT
   (tgBasicActuator.cpp:139-142); a NaN entry leaves that cable untouched. */
tgb_status tgb_set_control_input(tgb_world* w, int32_t env_begin, int32_t env_count,
                                 const double* input, int32_t broadcast, double dt);

/* n_steps x tgSimulation::step(dt) for every environment. Asynchronous on the world's stream. dt<=0 ->
   TGB_ERR_INVALID_ARGUMENT (tgSimulation.cpp:197-200). */
tgb_status tgb_step(tgb_world* w, double dt, int32_t n_steps);
tgb_status tgb_sync(tgb_world* w);

/* The two halves of one tgSimulation::step(dt), for callers that run host-side observers between them exactly where the
   reference runs them: tgb_step_world = tgWorld::step (src/core/tgSimulation.cpp:205, src/core/tgWorld.cpp:119-130);
   tgb_step_models = the models' step (tgSimulation.cpp:208-211): device-side controller, motors, cable forces.
   tgb_step(w, dt, 1) == tgb_step_world(w, dt) followed by tgb_step_models(w, dt). */
tgb_status tgb_step_world(tgb_world* w, double dt);
tgb_status tgb_step_models(tgb_world* w, double dt);

/* Host-buffer round trip in one call: upload rod+cable state, step, download rod state + cable observations.
   This is the reference-facing end-to-end path (bench.py "e2e"). Buffers as in the set/get calls, all envs. */
tgb_status tgb_rollout_host(tgb_world* w, const double* rod_state_in, const double* cable_rest_prev_in,
                            double dt, int32_t n_steps, double* rod_state_out, double* cable_obs_out);

/* Simulated time of the world = sum of dt over all steps (tgSimView.cpp m_totalTime equivalent). */
tgb_status tgb_get_time(const tgb_world* w, double* out);

/* Per-env status flags: bit0 = non-finite state seen, bit1 = centre of mass left [height_lo, height_hi],
   bit2 = a CPG update needed more than max_steps right-hand-side evaluations ("Inefficient CPG Parameters"). */
tgb_status tgb_get_status(tgb_world* w, int32_t env_begin, int32_t env_count, int32_t* out);

/* scores: [env_count][2]: XZ distance of the mass-weighted COM from its position at tgb_finalize /
   tgb_mark_episode_start, and 0 (energy: history logging is off in every shipped config). */
tgb_status tgb_mark_episode_start(tgb_world* w);
tgb_status tgb_get_scores(tgb_world* w, int32_t env_begin, int32_t env_count, double* scores);

/* Device-resident score buffer (float[n_env][2]) for NCCL allgather without a host hop. */
tgb_status tgb_scores_device_ptr(tgb_world* w, void** ptr, size_t* bytes);

/* The same without synchronising: the refresh is queued on the world's stream (tgb_stream_handle); order the consumer after it with
   an event on that stream.  For pipelines that overlap the all-gather of one episode chunk with the next chunk's steps. */
tgb_status tgb_scores_refresh_async(tgb_world* w, void** device_ptr, size_t* bytes);

/* Timing helper: run `reps` launches of n_steps each, bracketed by CUDA events on the world's stream;
   ms_out[reps]. Used by bench.py because torch.cuda.Event only sees torch's current stream. */
tgb_status tgb_step_timed(tgb_world* w, double dt, int32_t n_steps, int32_t reps, float* ms_out);

/* Kernel launches issued by this world since creation (bench.py "gpu_launches"). */
tgb_status tgb_launch_count(const tgb_world* w, int64_t* out);

/* Launch geometry chosen for this world: lanes per environment, threads per block, blocks. */
tgb_status tgb_launch_geometry(const tgb_world* w, int32_t* lanes_per_env, int32_t* block, int32_t* grid);
/* Override lanes per env (1,2,4,8,16,32) before tgb_finalize; 0 = automatic. */
tgb_status tgb_set_lanes_per_env(tgb_world* w, int32_t lanes);

/* Contact between the bodies of an environment (ABI version 4).  The reference's world collides every pair of rigid bodies
   (src/core/tgWorldBulletPhysicsImpl.cpp:82-107: btAxisSweep3 + btDefaultCollisionConfiguration; rods are btCylinderShape,
   src/tgcreator/tgRodInfo.cpp:99-113): rods of a tensegrity that fold onto each other, vertebrae of a spine that meet.  Here every pair
   of bodies with at least one dynamic member is tested on every step (cylinders as the capsules inscribed in them, spheres; boxes
   take no part) and a pair within the contact breaking threshold gets rows that couple the two bodies in the contact solver.
   `cap` = rows of this kind kept per environment (default TGB_DEFAULT_PAIR_CONTACT_CAP, at most 4 per pair of bodies); 0 switches
   the contact between bodies off (bodies then pass through each other, as in ABI version <= 3); -1 keeps the test but generates no row
   (tgb_get_min_clearance then measures how far bodies would interpenetrate).  Before tgb_finalize. */
#define TGB_DEFAULT_PAIR_CONTACT_CAP 6
tgb_status tgb_set_pair_contact_cap(tgb_world* w, int32_t cap);

/* Persistent contact manifolds with warm starting (ABI version 4).  Bullet keeps a manifold point's applied impulses from step to step
   (btPersistentManifold) and its sequential-impulse solver starts every solve from them scaled by
   btContactSolverInfo::m_warmstartingFactor = 0.85 (SOLVER_USE_WARMSTARTING is on by default; the world of
   src/core/tgWorldBulletPhysicsImpl.cpp:134-146 leaves both alone).  Here a point keeps its identity through its slot -- (shape, rim
   point) against the ground, (obstacle, shape, sample) against an obstacle, (shape, shape, point) between two bodies -- and the
   impulses travel with the environment between launches.  0 = every solve starts from zero.  Changing the factor empties the
   manifolds.  Default TGB_DEFAULT_WARM_STARTING = 0 (off): NTRTsim's world solves contacts with btMLCPSolver(btDantzigSolver), a direct
   solve that does not depend on a starting point, and on the reference suite's one quantitative whole-path test (ICRA2015_test.cpp) an
   ensemble of perturbed runs gives the same error distribution with and without warm starting, while the
   persistent manifolds cost ~20 % of the step's throughput.  Pass 0.85 for the behaviour of Bullet's own sequential-impulse solver. */
#define TGB_DEFAULT_WARM_STARTING 0.0
tgb_status tgb_set_warm_starting(tgb_world* w, double factor);

/* out[env_count]: the smallest surface distance between the collision shapes (capsules, spheres) of two different bodies of the
   environment that any step has seen since tgb_finalize or the last call with reset != 0 (1e30: no pair has come near; measured for
   pairs whose bounding spheres overlap).  Negative = the shapes overlap by that much. */
tgb_status tgb_get_min_clearance(tgb_world* w, int32_t env_begin, int32_t env_count, double* out, int32_t reset);

/* ---- Contact cables (ABI version 6) --------------------------------------------------------------------------
   tgBulletContactSpringCable (src/core/tgBulletContactSpringCable.cpp:97-190) as built by tgBasicContactCableInfo
   (src/tgcreator/tgBasicContactCableInfo.cpp:86-131): a spring-cable between two permanent anchors that collides with the rigid bodies
   of its world and bends around them over *sliding* anchors (src/core/tgBulletSpringCableAnchor.cpp:89-334).  Every step, in the models
   half (tgBulletContactSpringCable::step, :97-125):
     updateManifolds (:196-365)   contact points between the cable's segments (cylinders of radius `thickness` between consecutive
                                  anchors) and every body; a point with negative distance proposes a sliding anchor at the point on the
                                  body's surface, contact normal pointing from the cable into the body;
     updateAnchorList (:367-482)  a proposal joins the anchor list behind its nearest past anchor unless it lies within `resolution` of
                                  a sliding neighbour on the same body, the cable would have to push there (a neighbour lies behind
                                  the contact plane), a same-body neighbour's normal opposes its own, or the cable would grow by more than
                                  2 x resolution over the previous step's length;
     pruneAnchors (:563-673)      anchors whose body no longer touches the cable go; the rest slide sideways along
                                  ((back - forward) x normal) until the cable is straight in the tangent plane (updateAnchorPositions,
                                  :484-561, accepted while a contact point of the body lies within 0.1 of the new position); anchors
                                  closer than resolution / 2 to a neighbour or with a neighbour behind the contact plane go; repeated
                                  until four passes in a row remove nothing;
     calculateAndApplyForce (:127-194)  tension k (sum of segment lengths - rest length) >= 0 plus the clamped damping term; the end
                                  anchors are pulled along their segments, a sliding anchor gets magnitude x (unit(forward - current)
                                  + unit(back - current)); impulses force x dt at the anchors -- they sum to zero
                                  (test_integration/MuscleNP/MuscleNP_test.cpp:108-112 pins the momentum).
   Bullet's ghost object and persistent manifolds are restated geometrically, the way the contact between bodies is (see
   tgb_set_pair_contact_cap): a cylinder counts as the capsule inscribed in it, a sphere as itself, boxes take no part; the "manifold" of
   (cable, body) is the set of closest-point pairs of this step's segments with the body's shapes that are nearer than Bullet's contact
   breaking threshold (0.02), evaluated at the poses of the cable's step (Bullet's are one world step old).  At most
   TGB_CCABLE_MAX_SLIDING sliding anchors per cable (further proposals are dropped) and 64 manifold points per cable and step.  The actuator is tgBasicActuator: its step is the
   cable's step (src/core/tgBasicActuator.cpp:227-241); control input arrives through tgb_set_contact_cable_control_input (the
   device-side controllers of tgb_set_controller drive the two-anchor cables only).  Worlds with contact cables run one launch of
   the step kernel plus one of the contact-cable kernel per step. */
#define TGB_CCABLE_MAX_SLIDING 12
/* per-cable observation: rest length, actual length (over all anchors), tension, velocity, number of sliding anchors */
#define TGB_CCABLE_OBS_DOUBLES 5
/* per-anchor record of tgb_get_contact_cable_anchors: body index, world position (3), contact normal in the world frame (3; the
   anchor keeps it in its body's frame, tgBulletSpringCableAnchor.cpp:38,56-57,202-214) */
#define TGB_CCABLE_ANCHOR_DOUBLES 7
typedef struct {
    int32_t body_a, body_b;
    double local_a[3], local_b[3]; /* the two permanent anchors in their bodies' frames */
    double stiffness, damping;
    double rest_length;    /* |p2 - p1| - pretension / k */
    double prev_length;    /* == rest_length (tgBulletSpringCable.cpp:47, the first-step damping artefact) */
    double thickness;      /* tgBulletContactSpringCable.h:89: 0.001 */
    double resolution;     /* tgBulletContactSpringCable.h:90: 0.1 */
    /* the tgBasicActuator around it (tgSpringCableActuator::Config): used by tgb_set_contact_cable_control_input */
    double max_tension, target_velocity, min_actual_length, min_rest_length;
} tgb_contact_cable_desc;

/* Append contact cables (body indices relative to the model added LAST with tgb_add_model_flat).  Before tgb_finalize. */
tgb_status tgb_add_contact_cables(tgb_world* w, const tgb_contact_cable_desc* cables, int32_t n_cables);
tgb_status tgb_num_contact_cables(const tgb_world* w, int32_t* out);
/* obs: [env_count][n_contact_cables][TGB_CCABLE_OBS_DOUBLES] */
tgb_status tgb_get_contact_cable_obs(tgb_world* w, int32_t env_begin, int32_t env_count, double* obs);
/* the sliding anchors of one contact cable, in order from anchor 1 to anchor 2: out[env_count][TGB_CCABLE_MAX_SLIDING][7], unused
   slots zero with body index -1 (tgSpringCable::getAnchors without the two permanent ends) */
tgb_status tgb_get_contact_cable_anchors(tgb_world* w, int32_t cable, int32_t env_begin, int32_t env_count, double* out);
/* Start the contact cables of the environments in the range afresh, as tgBasicContactCableInfo builds them: rest length and previous
   length from rest_prev [env_count][n_contact_cables][2] (NULL: the descriptors' own), preferred length = rest length, velocity 0, no
   sliding anchors -- the counterpart of tgb_set_cable_state for a new episode from tgb_set_rod_state. */
tgb_status tgb_reset_contact_cables(tgb_world* w, int32_t env_begin, int32_t env_count, const double* rest_prev);
/* tgBasicActuator::setControlInput(input, dt) (src/core/tgBasicActuator.cpp:149-168) on the contact cables of every environment in the
   range, once, now: preferred length = input, then moveMotors(dt) (:170-225) with the cable's actual length over all its anchors.
   input: [n_contact_cables] (broadcast != 0) or [env_count][n_contact_cables]; negative -> TGB_ERR_INVALID_ARGUMENT, NaN = leave alone. */
tgb_status tgb_set_contact_cable_control_input(tgb_world* w, int32_t env_begin, int32_t env_count, const double* input, int32_t broadcast,
                                               double dt);

/* Threads per CTA of the step kernel (32, 64, 96, 128; a multiple of lanes per env); 0 = automatic
   (also settable through the environment variable TGB_BLOCK_THREADS). */
tgb_status tgb_set_block_threads(tgb_world* w, int32_t threads);

/* Layout of the step kernel: 0 = one group of lanes per environment does bodies and cables (best when the device is full of warps),
   1 = CTAs of four warps that hold 32 / lanes environments -- warp 0 runs the bodies, all four share the cables, 4 x lanes per
   environment (shortens a step when the kernel is latency bound: a few thousand environments per GPU), -1 = automatic (default;
   the environment variable TGB_WIDE=0/1 overrides the automatic choice).  Lanes per env must be 4 or 8 and the controller not the
   CPG one, else the narrow layout is used.  Results are bit-identical in both layouts.  May be called at any time. */
tgb_status tgb_set_wide_layout(tgb_world* w, int32_t mode);

/* Tail split of long launches (worlds whose bodies collide with each other): environments that spent at least 1/8 of their last
   launch with contact rows between two bodies (rods leaning on each other -- they run the coupled contact solver on every step and
   a launch ends with its slowest warp) get a warp of their own on a second stream, the others keep the usual layout.  -1 =
   automatic (default: launches of >= 64 steps over >= 512 environments; TGB_TAIL_SPLIT=0 in the environment turns it off), 0 = off,
   1 = every launch with a threshold of one step (tests).  Results are bit-identical with and without.  No reference counterpart:
   Bullet steps one world at a time.  May be called at any time. */
tgb_status tgb_set_tail_split(tgb_world* w, int32_t mode);

/* Bytes of device state per environment (the HBM traffic of one launch is 2x this per env). */
tgb_status tgb_state_bytes_per_env(const tgb_world* w, size_t* out);

/* Height window of the status flag bit1: the check of BaseSpineCPGControl::onStep
   (src/examples/learningSpines/BaseSpineCPGControl.cpp:271-278: 1.0 < y_COM < 25). lo > hi disables it (default). */
tgb_status tgb_set_height_window(tgb_world* w, double lo, double hi);

/* The CUDA stream (cudaStream_t) all work of this world is issued on. */
tgb_status tgb_stream_handle(tgb_world* w, void** stream);

#ifdef __cplusplus
}
#endif
#endif /* NTRTSIM_B200_H */
