/*
 * tgb_host_capi.cpp -- C entry points over the C++ host API, so that Python (ctypes) tests and bench.py can
 * build the shipped example models through tgStructure / tgBuildSpec / tgStructureInfo and read the flattened
 * descriptors, or drive a whole tgSimulation on the device.  World settings per example follow the reference
 * apps: AppPrismModel.cpp:46-80, AppSUPERball.cpp:45-91, AppEscape.cpp:61-151, MotorTimestep_test.cpp:81-215.
 */
#include "tgb_examples.h"
#include "tgb_sensors.h"
#include "tgb_yaml.h"

#include <cstring>
#include <string>

#include <dlfcn.h>

namespace {
thread_local std::string g_error;

/* "ref:<shared library>@<example>": the world of <example>, but the model objects come from `tgModel* tgb_make_model(const char*
   which)` of the library -- tests/ref_models builds one from the reference's own, unchanged model sources compiled against this
   host API (ntrtsim_b200/host/compat), which is what "the existing tgModel / tgStructure / tgBuildSpec API stays in place" means. */
typedef tgModel* (*ModelFactory)(const char*);
thread_local ModelFactory g_factory = 0;
tgModel* fromFactory(const std::string& which) {
    tgModel* m = g_factory(which.c_str());
    if (!m) throw std::invalid_argument("the model library has no model for " + which);
    return m;
}
#define TGB_MODEL(which, expr) (g_factory ? fromFactory(which) : static_cast<tgModel*>(expr))

struct Example {
    tgWorld* world;
    tgSimView* view;
    tgSimulation* sim;
    tgModel* model; /* owned by sim */
    tgModel* obstacle; /* a second model of the same world (owned by sim), or NULL */
    double dt;
    Example() : world(0), view(0), sim(0), model(0), obstacle(0), dt(0.001) {}
    ~Example() {
        delete sim;
        delete view;
        delete world;
    }
};

Example* build(const std::string& name, int n_env = 1, int precision = TGB_F32) {
    if (name.compare(0, 4, "ref:") == 0) {
        const std::size_t at = name.rfind('@');
        if (at == std::string::npos) { g_error = "ref:<library>@<example> expected"; return 0; }
        const std::string lib = name.substr(4, at - 4), base = name.substr(at + 1);
        void* h = dlopen(lib.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h) { g_error = std::string("cannot load the model library: ") + dlerror(); return 0; }
        ModelFactory f = reinterpret_cast<ModelFactory>(dlsym(h, "tgb_make_model"));
        if (!f) { g_error = "the model library does not export tgb_make_model"; return 0; }
        g_factory = f;
        Example* e = build(base, n_env, precision);
        g_factory = 0;
        return e; /* the library stays loaded: the model's code lives in it */
    }
    Example* e = new Example();
    try {
        if (name == "3_prism" || name == "3_prism_noground") {
            tgGround* ground = name == "3_prism" ? static_cast<tgGround*>(new tgBoxGround(tgBoxGround::Config(btVector3(0.0, 0.0, 0.0))))
                                                 : static_cast<tgGround*>(new tgEmptyGround());
            e->world = new tgWorld(tgWorld::Config(981), ground);
            e->dt = 0.001;
            e->model = TGB_MODEL(name, new PrismModel());
        } else if (name == "ContactCableDemo") {
            /* MuscleNP_test.cpp:80-96: no gravity, tgEmptyGround, dt = 1e-3 */
            e->world = new tgWorld(tgWorld::Config(0.0), new tgEmptyGround());
            e->dt = 0.001;
            e->model = TGB_MODEL(name, new ContactCableDemo());
        } else if (name == "SUPERball" || name == "SUPERball_flat") {
            const double pitch = name == "SUPERball" ? M_PI / 15.0 : 0.0;
            e->world = new tgWorld(tgWorld::Config(98.1), new tgBoxGround(tgBoxGround::Config(btVector3(0.0, pitch, 0.0))));
            e->dt = 0.001;
            e->model = TGB_MODEL(name, new T6Model());
        } else if (name == "Escape_crater" || name == "Escape_crater_shallow") {
            /* AppEscape.cpp:59-97 complete: EscapeModel and, as a second model of the same simulation, tgCraterDeep at the origin */
            e->world = new tgWorld(tgWorld::Config(98.1),
                                   new tgBoxGround(tgBoxGround::Config(btVector3(0, 0, 0), 0.5, 0.0, btVector3(10000.0, 2, 10000.0), btVector3(0, 0, 0))));
            e->dt = 1.0 / 60.0 / 10.0;
            e->model = TGB_MODEL(name, new EscapeModel());
            e->obstacle = TGB_MODEL(name + ":obstacle", name == "Escape_crater" ? static_cast<tgModel*>(new tgCraterDeep(btVector3(0, 0, 0)))
                                                                                : static_cast<tgModel*>(new tgCraterShallow(btVector3(0, 0, 0))));
        } else if (name == "SUPERball_stairs" || name == "SUPERball_blocks" || name == "SUPERball_wall") {
            /* src/models/obstacles/AppObstacleTest.cpp-style scenes: the SUPERball on the default ground next to an obstacle model */
            e->world = new tgWorld(tgWorld::Config(98.1));
            e->dt = 0.001;
            e->model = TGB_MODEL(name, new T6Model());
            if (name == "SUPERball_stairs") {
                e->obstacle = TGB_MODEL(name + ":obstacle", new tgStairs(tgStairs::Config(btVector3(0.0, 2.0, -10.0), 0.5, 0.0, 6, 20.0, 5.0, 1.0, 0.0)));
            } else if (name == "SUPERball_blocks") {
                srand(1); /* the reference draws the block positions from rand(); a fixed seed keeps the scene reproducible */
                e->obstacle = new tgBlockField(tgBlockField::Config(btVector3(0.0, 1.5, 0.0), 0.5, 0.0, btVector3(-30.0, 0.0, -30.0),
                                                                    btVector3(30.0, 0.0, 30.0), 20, 5.0, 5.0, 2.0));
            } else {
                e->obstacle = new Wall(btVector3(0.0, 2.5, 6.0));
            }
        } else if (name == "Escape" || name == "Escape_c3") {
            /* AppEscape: box ground half extents (10000, 2, 10000), dt = 1/600. "_c3" = BASELINE config 3:
               default box ground, dt = 1e-3. */
            tgBoxGround::Config gc = name == "Escape"
                                         ? tgBoxGround::Config(btVector3(0, 0, 0), 0.5, 0.0, btVector3(10000.0, 2, 10000.0), btVector3(0, 0, 0))
                                         : tgBoxGround::Config();
            e->world = new tgWorld(tgWorld::Config(98.1), new tgBoxGround(gc));
            e->dt = name == "Escape" ? 1.0 / 60.0 / 10.0 : 0.001;
            e->model = TGB_MODEL(name, new EscapeModel());
        } else if (name == "tsTestRig_basic" || name == "tsTestRig_kinematic") {
            e->world = new tgWorld(tgWorld::Config(981)); /* MotorTimestep_test: default ground */
            e->dt = 0.001;
            e->model = TGB_MODEL(name, new tsTestRig(name == "tsTestRig_kinematic"));
        } else if (name == "TetraSpine") {
            /* AppTetraSpineLearning.cpp:48-109: tgWorld(Config(981)) with the default box ground, dt = 1e-3, 3 segments */
            e->world = new tgWorld(tgWorld::Config(981));
            e->dt = 0.001;
            e->model = TGB_MODEL(name, new TetraSpineLearningModel(3));
        } else if (name == "TetraSpineStatic_hf") {
            /* ICRA2015_test.cpp:77-101: tgWorld(Config(981)) with the default box ground, dt = 1e-3, 3 segments */
            e->world = new tgWorld(tgWorld::Config(981));
            e->dt = 0.001;
            e->model = new TetraSpineStaticModel_hf(3);
        } else if (name == "FlemonsSpine" || name == "FlemonsSpine3") {
            /* WorldConf_Spines_test.cpp:85-108 / AppFlemonsSpineLearning: tgWorld(Config(981)), default box ground, dt = 1e-3, 12
               segments ("FlemonsSpine3": the same vertebrae, 3 segments, for small tests) */
            e->world = new tgWorld(tgWorld::Config(981));
            e->dt = 0.001;
            e->model = TGB_MODEL(name, new FlemonsSpineModelLearning(name == "FlemonsSpine" ? 12 : 3));
        } else if (name.size() > 5 && name.compare(0, 5, "yaml:") == 0) {
            /* "yaml:<path>[@gravity[@dt]]": a structure file of the YAML front-end in the world of the reference's yamlbuilder app
               (src/yamlbuilder/BuildTensegrityModel.cpp:44-79: gravity 98.1, default box ground, dt = 1e-3) */
            std::string path = name.substr(5);
            double gravity = 98.1, dt = 0.001;
            std::size_t at = path.find('@');
            if (at != std::string::npos) {
                std::string rest = path.substr(at + 1);
                path = path.substr(0, at);
                gravity = std::atof(rest.c_str());
                at = rest.find('@');
                if (at != std::string::npos) dt = std::atof(rest.substr(at + 1).c_str());
            }
            e->world = new tgWorld(tgWorld::Config(gravity));
            e->dt = dt;
            e->model = new TensegrityModel(path);
        } else {
            g_error = "unknown example: " + name;
            delete e;
            return 0;
        }
        e->world->setBatch(n_env, (tgb_precision)precision, 0);
        e->view = new tgSimView(*e->world, e->dt, 1.0 / 60.0 < e->dt ? e->dt : 1.0 / 60.0);
        e->sim = new tgSimulation(*e->view);
        e->sim->addModel(e->model);
        if (e->obstacle) e->sim->addModel(e->obstacle);
    } catch (const std::exception& ex) {
        g_error = ex.what();
        delete e;
        return 0;
    }
    return e;
}
} // namespace

extern "C" {

typedef struct tgh_model tgh_model;

const char* tgh_last_error(void) { return g_error.c_str(); }

tgh_model* tgh_build_example(const char* name) { return reinterpret_cast<tgh_model*>(build(name ? name : "")); }

void tgh_free(tgh_model* m) { delete reinterpret_cast<Example*>(m); }

int tgh_counts(const tgh_model* m, int* n_bodies, int* n_shapes, int* n_cables) {
    const Example* e = reinterpret_cast<const Example*>(m);
    if (!e) return 1;
    const tgWorldB200Impl& w = e->world->b200();
    *n_bodies = (int)w.bodies().size();
    *n_shapes = (int)w.shapes().size();
    *n_cables = (int)w.cables().size();
    return 0;
}

int tgh_copy(const tgh_model* m, tgb_body_desc* bodies, tgb_shape_desc* shapes, tgb_cable_desc* cables, tgb_ground_desc* ground, double* gravity,
             double* dt) {
    const Example* e = reinterpret_cast<const Example*>(m);
    if (!e) return 1;
    const tgWorldB200Impl& w = e->world->b200();
    if (bodies && !w.bodies().empty()) std::memcpy(bodies, w.bodies().data(), w.bodies().size() * sizeof(tgb_body_desc));
    if (shapes && !w.shapes().empty()) std::memcpy(shapes, w.shapes().data(), w.shapes().size() * sizeof(tgb_shape_desc));
    if (cables && !w.cables().empty()) std::memcpy(cables, w.cables().data(), w.cables().size() * sizeof(tgb_cable_desc));
    if (ground) *ground = w.ground();
    if (gravity) *gravity = w.gravity();
    if (dt) *dt = e->dt;
    return 0;
}

/* the contact cables of the built world: returns their number, copies at most `cap` */
int tgh_contact_cables(const tgh_model* m, tgb_contact_cable_desc* out, int cap) {
    const Example* e = reinterpret_cast<const Example*>(m);
    if (!e) return -1;
    const std::vector<tgb_contact_cable_desc>& cc = e->world->b200().contactCables();
    for (int i = 0; i < (int)cc.size() && i < cap && out; ++i) out[i] = cc[i];
    return (int)cc.size();
}

/* The model tree as the reference would print it: number of tgRod and tgSpringCableActuator descendants. */
int tgh_tree_counts(const tgh_model* m, int* n_rods, int* n_actuators) {
    const Example* e = reinterpret_cast<const Example*>(m);
    if (!e) return 1;
    std::vector<tgModel*> all = e->model->getDescendants();
    *n_rods = (int)tgCast::filter<tgModel, tgRod>(all).size();
    *n_actuators = (int)tgCast::filter<tgModel, tgSpringCableActuator>(all).size();
    return 0;
}

void tgh_escape_transform_actions(const double* actions01, double initial_length, double* out33) {
    tgbEscapeTransformActions(actions01, initial_length, out33);
}

/* The same, but stepping through tgSimulation::step one step at a time: the host model tree walk (observers, the models' own
   step code such as tsTestRig::step commanding its motor) runs between the world half and the models half of every step. */
int tgh_run_example_stepwise(const char* name, int n_env, int precision, int steps, double* rod_state_out, double* cable_obs_out) {
    Example* e = 0;
    try {
        e = build(name ? name : "", n_env, precision);
        if (!e) return TGB_ERR_INVALID_ARGUMENT;
        for (int i = 0; i < steps; ++i) e->sim->step(e->dt);
        tgWorldB200Impl& w = e->world->b200();
        int st = TGB_OK;
        if (rod_state_out) st = tgb_get_rod_state(w.handle(), 0, n_env, rod_state_out);
        if (st == TGB_OK && cable_obs_out) st = tgb_get_cable_obs(w.handle(), 0, n_env, cable_obs_out);
        delete e;
        return st;
    } catch (const std::invalid_argument& ex) {
        g_error = ex.what();
        delete e;
        return TGB_ERR_INVALID_ARGUMENT;
    } catch (const std::exception& ex) {
        g_error = ex.what();
        delete e;
        return TGB_ERR_RUNTIME;
    }
}

/* An example app with a tgDataLogger2 attached the way the reference apps do it (e.g. src/examples/3_prism/AppPrismModel.cpp with
   tgDataLogger2 + tgRodSensorInfo + tgSpringCableActuatorSensorInfo): `steps` steps through tgSimulation::run, one CSV line per
   sample of environment `env`.  The name of the written file is returned through file_name_out. */
int tgh_run_example_logged(const char* name, int n_env, int precision, int steps, const char* log_prefix, double time_interval, int env,
                           char* file_name_out, int file_name_cap) {
    Example* e = 0;
    try {
        e = build(name ? name : "", n_env, precision);
        if (!e) return TGB_ERR_INVALID_ARGUMENT;
        tgDataLogger2* logger = new tgDataLogger2(log_prefix ? log_prefix : "", time_interval, env);
        logger->addSenseable(e->model);
        logger->addSensorInfo(new tgRodSensorInfo());
        logger->addSensorInfo(new tgSpringCableActuatorSensorInfo());
        logger->addSensorInfo(new tgCompoundRigidSensorInfo()); /* as in src/examples/.../AppHorizontalSpine.cpp */
        e->sim->addDataManager(logger);
        e->sim->run(steps);
        if (file_name_out && file_name_cap > 0) {
            std::strncpy(file_name_out, logger->getFileName().c_str(), (size_t)file_name_cap - 1);
            file_name_out[file_name_cap - 1] = 0;
        }
        delete e;
        return TGB_OK;
    } catch (const std::invalid_argument& ex) {
        g_error = ex.what();
        delete e;
        return TGB_ERR_INVALID_ARGUMENT;
    } catch (const std::exception& ex) {
        g_error = ex.what();
        delete e;
        return TGB_ERR_RUNTIME;
    }
}

/* CPG + impedance controller of TetraSpineCPGControl flattened for tgb_set_cpg_controller
   (AppTetraSpineLearning.cpp:77-86: Config(segmentSpan 3, 6, 6, 2, segment 1, controlTime .001);
   TetraSpineCPGControl.cpp:58-126: impedance gains by the "outer" tag).
   Call with NULL arrays to get the sizes: n_nodes, n_couplings, n_pairs. */
int tgh_tetraspine_cpg(const tgh_model* m, int* n_nodes, int* n_couplings, int* n_pairs, int32_t* begin, int32_t* target, int32_t* pair,
                       double* impedance, double* scalars /* r_const, d_min, d_max, desc, control_time, cable_step, lo, hi */,
                       int* ints /* max_steps, score_body */) {
    const Example* e = reinterpret_cast<const Example*>(m);
    if (!e) return 1;
    const tgWorldB200Impl& w = e->world->b200();
    const tgbSpineControlConfig cfg(3, 6, 6, 2, 1, .001);
    std::vector<int32_t> b, t, p;
    try {
        tgbSpineCpgTopology(w.cables(), cfg, b, t, p);
    } catch (const std::exception& ex) {
        g_error = ex.what();
        return 5;
    }
    *n_nodes = (int)w.cables().size();
    *n_couplings = (int)t.size();
    *n_pairs = cfg.tablePairs();
    if (begin) std::memcpy(begin, b.data(), b.size() * sizeof(int32_t));
    if (target) std::memcpy(target, t.data(), t.size() * sizeof(int32_t));
    if (pair) std::memcpy(pair, p.data(), p.size() * sizeof(int32_t));
    if (impedance) {
        const TetraSpineLearningModel* model = dynamic_cast<const TetraSpineLearningModel*>(e->model);
        if (!model) return 1;
        const std::vector<tgSpringCableActuator*>& muscles = model->getAllMuscles();
        for (std::size_t i = 0; i < muscles.size(); ++i) {
            const bool outer = muscles[i]->hasTag("outer");
            double* g = impedance + 4 * muscles[i]->cableIndex();
            g[0] = 100.0; g[1] = outer ? 100.0 : 500.0; g[2] = 200.0; g[3] = outer ? 19.5 : 16.5;
        }
    }
    if (scalars) {
        scalars[0] = 20.0; scalars[1] = 0.0; scalars[2] = 5.0; scalars[3] = 2.0; scalars[4] = cfg.controlTime; scalars[5] = 1.0 / 10000.0;
        scalars[6] = 1.0; scalars[7] = 25.0;
    }
    if (ints) { ints[0] = 200; ints[1] = cfg.segmentNumber; }
    return 0;
}

/* The controller of the kinematic tetrahedral-complex spine, flattened for tgb_set_cpg_controller with cable_control = 1:
   KinematicSpineCPGControl with the Config of WorldConf_Spines_test.cpp:110-147 (segmentSpan 3, 8 muscles, 2 params, scored segment 6 --
   clamped to the last segment of a shorter spine --, controlTime .01, amplitude and frequency ranges [-30, 30], impedance gains
   kt 0 / kp 1000 / kv 210 with the default control length = each muscle's start length), tgCPGCableControl(tgPIDController::Config(5000,
   0, 10, true), controlTime) per muscle (KinematicSpineCPGControl.cpp:69-76).  scalars additionally carry the PID gains [8..10]. */
static tgbSpineControlConfig flemonsControlConfig(std::size_t segments) {
    tgbSpineControlConfig cfg(3, 8, 8, 2, segments > 6 ? 6 : (int)segments - 1, .01);
    cfg.lowAmp = -30.0; cfg.highAmp = 30.0; cfg.lowFreq = -30.0; cfg.highFreq = 30.0;
    return cfg;
}
int tgh_flemons_cpg(const tgh_model* m, int* n_nodes, int* n_couplings, int* n_pairs, int32_t* begin, int32_t* target, int32_t* pair,
                    double* impedance, double* scalars /* r_const, d_min, d_max, desc, control_time, cable_step, lo, hi, kP, kI, kD */,
                    int* ints /* max_steps, score_body */) {
    const Example* e = reinterpret_cast<const Example*>(m);
    if (!e) return 1;
    const FlemonsSpineModelLearning* model = dynamic_cast<const FlemonsSpineModelLearning*>(e->model);
    if (!model) { g_error = "not a FlemonsSpineModelLearning"; return 1; }
    const tgWorldB200Impl& w = e->world->b200();
    const tgbSpineControlConfig cfg = flemonsControlConfig(model->getSegments());
    std::vector<int32_t> b, t, p;
    try {
        tgbSpineCpgTopology(w.cables(), cfg, b, t, p);
    } catch (const std::exception& ex) {
        g_error = ex.what();
        return 5;
    }
    *n_nodes = (int)w.cables().size();
    *n_couplings = (int)t.size();
    *n_pairs = cfg.tablePairs();
    if (begin) std::memcpy(begin, b.data(), b.size() * sizeof(int32_t));
    if (target) std::memcpy(target, t.data(), t.size() * sizeof(int32_t));
    if (pair) std::memcpy(pair, p.data(), p.size() * sizeof(int32_t));
    if (impedance) {
        const std::vector<tgSpringCableActuator*>& muscles = model->getAllMuscles();
        for (std::size_t i = 0; i < muscles.size(); ++i) {
            double* g = impedance + 4 * muscles[i]->cableIndex();
            g[0] = 0.0; g[1] = 1000.0; g[2] = 210.0;
            g[3] = muscles[i]->getStartLength(); /* tgCPGActuatorControl::onAttach, tgCPGActuatorControl.cpp:64-66 */
        }
    }
    if (scalars) {
        scalars[0] = 20.0; scalars[1] = 0.0; scalars[2] = 5.0; scalars[3] = 2.0; scalars[4] = cfg.controlTime; scalars[5] = cfg.controlTime;
        scalars[6] = 1.0; scalars[7] = 25.0;
        scalars[8] = -5000.0; scalars[9] = -0.0; scalars[10] = -10.0; /* tension control inverts the gains, tgPIDController.cpp:34-43 */
    }
    if (ints) { ints[0] = 200; ints[1] = cfg.segmentNumber; }
    return 0;
}
int tgh_flemons_param_row(const tgh_model* m, const double* node_actions, const double* edge_actions, int n_edge, double* row) {
    const Example* e = reinterpret_cast<const Example*>(m);
    const FlemonsSpineModelLearning* model = e ? dynamic_cast<const FlemonsSpineModelLearning*>(e->model) : 0;
    if (!model) return -1;
    const tgbSpineControlConfig cfg = flemonsControlConfig(model->getSegments());
    tgbSpineScaleNodeActions(node_actions, cfg, row);
    return tgbSpineScaleEdgeActions(edge_actions, n_edge, cfg, row + 2);
}

/* scaleNodeActions + scaleEdgeActions for the TetraSpine configuration: node_actions[2], edge_actions[n_edge][2] in [0,1]
   -> one parameter row [2 + 2 * n_pairs] for tgb_set_cpg_controller. Returns the number of edge actions consumed. */
int tgh_tetraspine_param_row(const double* node_actions, const double* edge_actions, int n_edge, double* row) {
    const tgbSpineControlConfig cfg(3, 6, 6, 2, 1, .001);
    tgbSpineScaleNodeActions(node_actions, cfg, row);
    return tgbSpineScaleEdgeActions(edge_actions, n_edge, cfg, row + 2);
}

/* tgUtil::getQuaternionBetween (src/tgcreator/tgUtil.h:199-238) and btVector3::rotate, for the test that mirrors
   test/tgcreator/tgUtil_test.cpp:72-97.  q_out = (x, y, z, w); rotated_out = v rotated by that quaternion's axis and angle. */
void tgh_quaternion_between(const double* a, const double* b, const double* v, double* q_out, double* rotated_out) {
    const btQuaternion q = tgUtil::getQuaternionBetween(btVector3(a[0], a[1], a[2]), btVector3(b[0], b[1], b[2]));
    q_out[0] = q.x(); q_out[1] = q.y(); q_out[2] = q.z(); q_out[3] = q.w();
    const btVector3 r = btVector3(v[0], v[1], v[2]).rotate(q.getAxis(), q.getAngle());
    rotated_out[0] = r.x(); rotated_out[1] = r.y(); rotated_out[2] = r.z();
}

/* The YAML reader's view of a document as JSON (all scalars as strings, null as null): lets the tests compare the reader with a full
   YAML parser on the structure-file subset.  Returns the length needed (without the terminator), or -1 with tgh_last_error set. */
static void yamToJson(const Yam& y, std::string& out) {
    switch (y.type()) {
    case Yam::Null: out += "null"; break;
    case Yam::Scalar: {
        out += '"';
        for (std::size_t i = 0; i < y.str().size(); ++i) {
            const char c = y.str()[i];
            if (c == '"' || c == '\\') out += '\\';
            out += c;
        }
        out += '"';
        break;
    }
    case Yam::Sequence:
        out += '[';
        for (std::size_t i = 0; i < y.elements().size(); ++i) {
            if (i) out += ", ";
            yamToJson(y.elements()[i], out);
        }
        out += ']';
        break;
    case Yam::Map:
        out += '{';
        for (std::size_t i = 0; i < y.items().size(); ++i) {
            if (i) out += ", ";
            yamToJson(Yam::load("\"" + y.items()[i].first + "\""), out);
            out += ": ";
            yamToJson(y.items()[i].second, out);
        }
        out += '}';
        break;
    }
}
int tgh_yaml_to_json(const char* text, char* out, int cap) {
    try {
        std::string json;
        yamToJson(Yam::load(text ? text : ""), json);
        if (out && cap > 0) {
            std::strncpy(out, json.c_str(), (size_t)cap - 1);
            out[cap - 1] = 0;
        }
        return (int)json.size();
    } catch (const std::exception& ex) {
        g_error = ex.what();
        return -1;
    }
}

/* Drop-in demonstration: the example app's main loop on the device through the host API only
   (tgSimulation::run -> tgWorldB200Impl -> C ABI). Outputs like tgb_get_rod_state / tgb_get_cable_obs. */
int tgh_run_example(const char* name, int n_env, int precision, int steps, double* rod_state_out, double* cable_obs_out) {
    Example* e = 0;
    try {
        e = build(name ? name : "", n_env, precision);
        if (!e) return TGB_ERR_INVALID_ARGUMENT;
        e->sim->run(steps);
        tgWorldB200Impl& w = e->world->b200();
        int st = TGB_OK;
        if (rod_state_out) st = tgb_get_rod_state(w.handle(), 0, n_env, rod_state_out);
        if (st == TGB_OK && cable_obs_out) st = tgb_get_cable_obs(w.handle(), 0, n_env, cable_obs_out);
        delete e;
        return st;
    } catch (const std::invalid_argument& ex) {
        g_error = ex.what();
        delete e;
        return TGB_ERR_INVALID_ARGUMENT;
    } catch (const std::exception& ex) {
        g_error = ex.what();
        delete e;
        return TGB_ERR_RUNTIME;
    }
}

} // extern "C"
