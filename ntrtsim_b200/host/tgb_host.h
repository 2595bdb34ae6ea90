/*
 * tgb_host.h -- host side of the drop-in boundary: the NTRTsim modelling / simulation API
 * (tgStructure, tgBuildSpec, tgModel, tgSpringCableActuator, tgObserver, tgWorld, tgSimulation ...)
 * re-implemented over the batched B200 world.  Models are described exactly as with the reference
 * (nodes, pairs, tags, build spec); buildInto() flattens them into the descriptors of
 * include/ntrtsim_b200.h instead of creating Bullet objects, and tgSimulation::step / run drive the
 * device through the C ABI.  Class and member names, argument meaning and exception types follow the
 * reference headers cited at each class; the implementation is new.
 *
 * Batched extension (not in the reference): a tgWorld holds `n_env` independent copies of the model
 * (tgWorld::setBatch).  Getters take an optional environment index, default 0.
 */
#ifndef TGB_HOST_H
#define TGB_HOST_H

#include "../../include/ntrtsim_b200.h"
#include "LinearMath/btTransform.h"

/* the standard headers the reference's core headers bring along (src/core/tgTaggable.h:29-35, tgModel.h:35-36): model sources
   written against them rely on assert, atoi, std::cout ... being declared */
#include <assert.h>
#include <stdlib.h>
#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <deque>
#include <iostream>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

/* ------------------------------------------------------------------------------------------------
 * Exceptions and tags  (src/core/tgException.h, tgTags.h, tgTaggable.h, tgTagSearch.h)
 * ---------------------------------------------------------------------------------------------- */
class tgException : public std::runtime_error {
public:
    explicit tgException(const std::string& what) : std::runtime_error(what) {}
};

/* A set of whitespace-separated words. */
class tgTags {
public:
    tgTags() {}
    tgTags(const std::string& space_separated) { append(space_separated); }
    tgTags(const char* space_separated) { append(std::string(space_separated ? space_separated : "")); }
    void append(const std::string& space_separated);
    void append(const tgTags& other);
    void remove(const tgTags& other);
    bool contains(const std::string& word) const;
    bool containsAll(const tgTags& other) const;
    bool containsAny(const tgTags& other) const;
    bool empty() const { return m_words.empty(); }
    std::size_t size() const { return m_words.size(); }
    const std::deque<std::string>& getTags() const { return m_words; }
    std::string joinTags(const std::string& sep = " ") const;
    tgTags operator+(const tgTags& o) const { tgTags r(*this); r.append(o); return r; }
    tgTags& operator+=(const tgTags& o) { append(o); return *this; }
private:
    std::deque<std::string> m_words;
};
std::ostream& operator<<(std::ostream& os, const tgTags& t);

/* "a b" = both, "a|b" = either, "-a" = must not have a. */
class tgTagSearch {
public:
    tgTagSearch() {}
    tgTagSearch(const std::string& expr); /* the only converting constructor, as in src/core/tgTagSearch.h:54: find<T>("tag") then picks
                                             the std::string overload of tgModel::find instead of being ambiguous */
    bool matches(const tgTags& tags) const;
    void remove(const tgTags& tags);
private:
    struct Alternative { tgTags required, forbidden; };
    std::vector<Alternative> m_alternatives;
    void parse(const std::string& expr);
};

class tgTaggable {
public:
    tgTaggable() {}
    tgTaggable(const tgTags& tags) : m_tags(tags) {}
    virtual ~tgTaggable() {}
    void addTags(const std::string& t) { m_tags.append(t); }
    void addTags(const tgTags& t) { m_tags.append(t); }
    void setTags(const tgTags& t) { m_tags = t; }
    bool hasTag(const std::string& t) const { return m_tags.contains(t); }
    bool hasAllTags(const std::string& t) const { return m_tags.containsAll(tgTags(t)); }
    bool hasAnyTags(const std::string& t) const { return m_tags.containsAny(tgTags(t)); }
    bool hasNoTags() const { return m_tags.empty(); }
    const tgTags& getTags() const { return m_tags; }
    tgTags& getTags() { return m_tags; }
    std::string getTagStr(const std::string& sep = " ") const { return m_tags.joinTags(sep); }
private:
    tgTags m_tags;
};

/* ------------------------------------------------------------------------------------------------
 * Observer pattern  (src/core/tgObserver.h:34-69, tgSubject.h:43-131)
 * ---------------------------------------------------------------------------------------------- */
template <class Subject>
class tgObserver {
public:
    virtual ~tgObserver() {}
    virtual void onStep(Subject& subject, double dt) = 0;
    virtual void onAttach(Subject& subject) { (void)subject; }
    virtual void onSetup(Subject& subject) { (void)subject; }
    virtual void onTeardown(Subject& subject) { (void)subject; }
};

template <class Subject>
class tgSubject {
public:
    virtual ~tgSubject() {}
    /* Observers are not owned. */
    void attach(tgObserver<Subject>* o) {
        if (!o) return;
        m_observers.push_back(o);
        o->onAttach(static_cast<Subject&>(*this));
    }
    void notifyStep(double dt) {
        if (!(dt > 0)) return;
        for (std::size_t i = 0; i < m_observers.size(); ++i) m_observers[i]->onStep(static_cast<Subject&>(*this), dt);
    }
    void notifySetup() {
        for (std::size_t i = 0; i < m_observers.size(); ++i) m_observers[i]->onSetup(static_cast<Subject&>(*this));
    }
    void notifyTeardown() {
        for (std::size_t i = 0; i < m_observers.size(); ++i) m_observers[i]->onTeardown(static_cast<Subject&>(*this));
    }
    std::size_t numObservers() const { return m_observers.size(); }
private:
    std::vector<tgObserver<Subject>*> m_observers;
};

/* ------------------------------------------------------------------------------------------------
 * Ground and world  (src/core/terrain/*.h, src/core/tgWorld.h:38-133, tgWorldImpl.h:39-58)
 * ---------------------------------------------------------------------------------------------- */
class tgGround {
public:
    virtual ~tgGround() {}
    /* The contact surface this ground presents to the batched step. */
    virtual tgb_ground_desc describe() const = 0;
};

class tgEmptyGround : public tgGround {
public:
    tgb_ground_desc describe() const;
};

/* Static box; bodies rest on its top face (tgBoxGround.cpp:55-100). As in the reference the Config's friction and
   restitution are not applied to the body: the surface has Bullet's defaults 0.5 / 0. */
class tgBoxGround : public tgGround {
public:
    struct Config {
        Config(btVector3 eulerAngles = btVector3(0.0, 0.0, 0.0), btScalar friction = 0.5, btScalar restitution = 0.0,
               btVector3 size = btVector3(500.0, 1.5, 500.0), btVector3 origin = btVector3(0.0, 0.0, 0.0));
        btVector3 m_eulerAngles; /* yaw, pitch, roll */
        btScalar m_friction;
        btScalar m_restitution;
        btVector3 m_size;        /* half extents */
        btVector3 m_origin;
    };
    tgBoxGround();
    explicit tgBoxGround(const Config& config);
    tgb_ground_desc describe() const;
private:
    Config m_config;
};

/* Infinite plane through the origin region (tgPlaneGround.cpp:52-96); friction / restitution ARE applied. */
class tgPlaneGround : public tgGround {
public:
    struct Config {
        Config(btVector3 normalVector = btVector3(0.0, 1.0, 0.0), btScalar friction = 0.5, btScalar restitution = 0.0,
               btVector3 origin = btVector3(0.0, 0.0, 0.0));
        btVector3 m_normalVector;
        btScalar m_friction;
        btScalar m_restitution;
        btVector3 m_origin;
    };
    tgPlaneGround();
    explicit tgPlaneGround(const Config& config);
    tgb_ground_desc describe() const;
private:
    Config m_config;
};

class tgWorld;

/* The physics seam of the reference: ctor + step(dt). */
class tgWorldImpl {
public:
    virtual ~tgWorldImpl() {}
    virtual void step(double dt) = 0;
};

/* The batched replacement for tgWorldBulletPhysicsImpl. Collects flat descriptors while models build, then owns
   the device world. */
class tgWorldB200Impl : public tgWorldImpl {
public:
    tgWorldB200Impl(double gravity, const tgGround* ground);
    ~tgWorldB200Impl();

    /* builder side */
    int addBody(const tgb_body_desc& body, const std::vector<tgb_shape_desc>& shapes);
    int addCable(const tgb_cable_desc& cable);
    /* contact cables (tgBulletContactSpringCable) live in their own list; the returned index is kContactCableBase + position and is
       understood by the cable observation calls below */
    enum { kContactCableBase = 1 << 20 };
    int addContactCable(const tgb_contact_cable_desc& cable);
    const std::vector<tgb_contact_cable_desc>& contactCables() const { return m_contactCables; }
    int contactCableAnchors(int cable, int env); /* number of sliding anchors (tgSpringCable::getAnchors().size() - 2) */
    const std::vector<tgb_body_desc>& bodies() const { return m_bodies; }
    const std::vector<tgb_shape_desc>& shapes() const { return m_shapes; }
    const std::vector<tgb_cable_desc>& cables() const { return m_cables; }
    tgb_ground_desc ground() const { return m_ground; }
    double gravity() const { return m_gravity; }

    /* batch configuration (before the first step) */
    void setBatch(int n_env, tgb_precision precision = TGB_F32, int device = 0);
    int numEnvs() const { return m_nEnv; }

    /* device side */
    bool finalized() const { return m_handle != 0; }
    void finalize();                 /* throws std::runtime_error when no CUDA device is usable */
    tgb_world* handle();             /* finalizes on first use */
    void step(double dt);            /* one tgSimulation::step for every environment */
    void stepWorld(double dt);       /* its world half (tgSimulation.cpp:205) */
    void stepModels(double dt);      /* its models half: device-side controller, motors, cable forces (tgSimulation.cpp:208-211) */
    void setMotorTorque(int cable, double torque); /* tgKinematicActuator::setControlInput, all environments, until changed */
    void run(double dt, int steps);  /* `steps` of them in one device launch sequence */

    /* observation side: valid for the state after the last completed step */
    double cableRestLength(int cable, int env);
    double cableLength(int cable, int env);
    double cableTension(int cable, int env);
    double cableMotorVelocity(int cable, int env);
    btVector3 bodyPosition(int body, int env);
    btQuaternion bodyOrientation(int body, int env);
    void setControlInput(int cable, double input, double dt); /* all environments */
    void invalidateCache() { m_cacheValid = false; }

private:
    void check(int status) const;
    void refreshCache();
    double m_gravity;
    tgb_ground_desc m_ground;
    std::vector<tgb_body_desc> m_bodies;
    std::vector<tgb_shape_desc> m_shapes;
    std::vector<tgb_cable_desc> m_cables;
    std::vector<tgb_contact_cable_desc> m_contactCables;
    std::vector<double> m_ccCache;
    int m_nEnv;
    tgb_precision m_precision;
    int m_device;
    tgb_world* m_handle;
    bool m_cacheValid;
    std::vector<double> m_rodCache, m_cableCache;
    std::vector<double> m_torque;
};

class tgWorld {
public:
    struct Config {
        Config(double g = 9.81, double ws = 1000);
        double gravity;
        double worldSize;
    };
    tgWorld();
    explicit tgWorld(const Config& config);
    tgWorld(const Config& config, tgGround* ground); /* takes ownership of ground */
    ~tgWorld();
    void reset();
    void reset(const Config& config);
    void reset(tgGround* ground);
    void step(double dt) const;
    tgWorldImpl& implementation() const { return *m_pImpl; }
    tgWorldB200Impl& b200() const { return *m_pImpl; }
    double getWorldGravity() const { return m_config.gravity; }
    /* batched extension */
    void setBatch(int n_env, tgb_precision precision = TGB_F32, int device = 0);
private:
    Config m_config;
    tgGround* m_pGround;
    tgWorldB200Impl* m_pImpl;
    int m_nEnv;
    tgb_precision m_precision;
    int m_device;
};

/* ------------------------------------------------------------------------------------------------
 * Model tree  (src/core/tgModel.h:49-164, tgCast.h)
 * ---------------------------------------------------------------------------------------------- */
class tgModelVisitor;

/* Anything a sensor can be attached to (src/core/tgSenseable.h:41-66). */
class tgSenseable {
public:
    virtual ~tgSenseable() {}
    virtual std::vector<tgSenseable*> getSenseableDescendants() const { return std::vector<tgSenseable*>(); }
};

class tgModel : public tgTaggable, public tgSenseable {
public:
    tgModel();
    explicit tgModel(const tgTags& tags);
    virtual ~tgModel();
    virtual void setup(tgWorld& world);
    virtual void teardown();
    virtual void step(double dt);
    /* batched extension: true when this model's step() does work on the host every step (its own control code or model-level
       observers), so that tgSimView::run must not fuse steps on the device */
    virtual bool stepsOnHost() const { return false; }
    virtual void onVisit(const tgModelVisitor& r) const;
    void addChild(tgModel* pChild); /* takes ownership; NULL -> std::invalid_argument */
    virtual std::string toString(std::string prefix = "") const;
    std::vector<tgModel*> getDescendants() const;
    virtual std::vector<tgSenseable*> getSenseableDescendants() const; /* tgModel.cpp:191-201 */
    const std::vector<tgModel*>& getChildren() const { return m_children; }
    template <typename T>
    std::vector<T*> find(const tgTagSearch& tagSearch) const {
        std::vector<T*> out;
        const std::vector<tgModel*> all = getDescendants();
        for (std::size_t i = 0; i < all.size(); ++i) {
            T* t = dynamic_cast<T*>(all[i]);
            if (t && tagSearch.matches(all[i]->getTags())) out.push_back(t);
        }
        return out;
    }
    template <typename T>
    std::vector<T*> find(const std::string& tagSearch) const { return find<T>(tgTagSearch(tagSearch)); }
private:
    std::vector<tgModel*> m_children;
};

/* tgModel.h:227-237: a model prints as its toString() */
inline std::ostream& operator<<(std::ostream& os, const tgModel& obj) { return os << obj.toString(); }

class tgModelVisitor {
public:
    virtual ~tgModelVisitor() {}
    virtual void render(const tgModel& model) const { (void)model; }
};

struct tgCast {
    template <typename T_FROM, typename T_TO>
    static std::vector<T_TO*> filter(const std::vector<T_FROM*>& v) {
        std::vector<T_TO*> out;
        for (std::size_t i = 0; i < v.size(); ++i) {
            T_TO* t = dynamic_cast<T_TO*>(v[i]);
            if (t) out.push_back(t);
        }
        return out;
    }
    template <typename T_FROM, typename T_TO>
    static T_TO* cast(T_FROM* obj) { return dynamic_cast<T_TO*>(obj); }
    template <typename T_FROM, typename T_TO>
    static std::vector<T_TO*> find(const tgTagSearch& s, const std::vector<T_FROM*>& v) {
        std::vector<T_TO*> out;
        for (std::size_t i = 0; i < v.size(); ++i) {
            T_TO* t = dynamic_cast<T_TO*>(v[i]);
            if (t && s.matches(v[i]->getTags())) out.push_back(t);
        }
        return out;
    }
};

/* ------------------------------------------------------------------------------------------------
 * Rigid bodies  (src/core/tgBaseRigid.h, tgRod.h:52-85)
 * ---------------------------------------------------------------------------------------------- */
class tgBaseRigid : public tgModel {
public:
    virtual double mass() const { return m_mass; }
    virtual btVector3 centerOfMass() const { return centerOfMass(0); }
    btVector3 centerOfMass(int env) const;
    /* yaw, pitch, roll of the body (tgBaseRigid.cpp:92-103) */
    virtual btVector3 orientation() const { return orientation(0); }
    btVector3 orientation(int env) const;
    btQuaternion orientationQuaternion(int env = 0) const; /* btRigidBody::getOrientation */
    int bodyIndex() const { return m_body; }
protected:
    tgBaseRigid(tgWorldB200Impl* world, int body, double mass, const tgTags& tags);
    tgWorldB200Impl* m_world;
    int m_body;
    double m_mass;
};

class tgRod : public tgBaseRigid {
public:
    struct Config {
        Config(double r = 0.5, double d = 1.0, double f = 1.0, double rf = 0.0, double res = 0.2);
        const double radius;
        const double density;
        const double friction;
        const double rollFriction;
        const double restitution;
    };
    tgRod(tgWorldB200Impl* world, int body, double mass, const tgTags& tags, double length);
    double length() const { return m_length; }
private:
    double m_length;
};

/* A box between two points (src/core/tgBox.h:43-118, tgBox.cpp:36-64): `width` and `height` are HALF extents across the axis */
class tgBox : public tgBaseRigid {
public:
    struct Config {
        Config(double w = 1.0, double h = 1.0, double d = 1.0, double f = 1.0, double rf = 0.0, double res = 0.2);
        const double width;
        const double height;
        const double density;
        const double friction;
        const double rollFriction;
        const double restitution;
    };
    tgBox(tgWorldB200Impl* world, int body, double mass, const tgTags& tags, double length);
    double length() const { return m_length; }
private:
    double m_length;
};

/* A sphere at a node (src/core/tgSphere.h:52-88, tgSphere.cpp:36-72) */
class tgSphere : public tgBaseRigid {
public:
    struct Config {
        Config(double r = 0.5, double d = 1.0, double f = 1.0, double rf = 0.0, double res = 0.2);
        const double radius;
        const double density;
        const double friction;
        const double rollFriction;
        const double restitution;
    };
    tgSphere(tgWorldB200Impl* world, int body, double mass, const tgTags& tags);
};

/* ------------------------------------------------------------------------------------------------
 * Cables and actuators  (src/core/tgSpringCableActuator.h:51-266, tgBasicActuator.h, tgKinematicActuator.h)
 * ---------------------------------------------------------------------------------------------- */
class tgControllable {
public:
    virtual ~tgControllable() {}
    virtual void setControlInput(double input) = 0;
    virtual void setControlInput(double input, double dt) { (void)dt; setControlInput(input); }
};

class tgSpringCableActuator : public tgModel, public tgControllable {
public:
    struct Config {
        Config(double s = 1000.0, double d = 10.0, double p = 0.0, bool h = false, double mf = 1000.0, double tVel = 100.0,
               double mnAL = 0.1, double mnRL = 0.1, double rot = 0, bool moveCPA = true, bool moveCPB = true);
        void scale(double sf);
        double stiffness;
        double damping;
        double pretension;
        bool hist;
        double maxTens;
        double targetVelocity;
        double minActualLength;
        double minRestLength;
        double rotation;
        bool moveCablePointAToEdge;
        bool moveCablePointBToEdge;
    };
    struct SpringCableActuatorHistory {
        std::deque<double> lastLengths, restLengths, dampingHistory, lastVelocities, tensionHistory;
    };
    virtual ~tgSpringCableActuator();
    virtual const double getStartLength() const { return m_startLength; }
    virtual const double getCurrentLength() const { return getCurrentLength(0); }
    virtual const double getTension() const { return getTension(0); }
    virtual const double getRestLength() const { return getRestLength(0); }
    virtual const double getVelocity() const { return 0.0; }
    double getCurrentLength(int env) const;
    double getTension(int env) const;
    double getRestLength(int env) const;
    virtual const SpringCableActuatorHistory& getHistory() const { return m_history; }
    const Config& getConfig() const { return m_config; }
    int cableIndex() const { return m_cable; }
protected:
    tgSpringCableActuator(tgWorldB200Impl* world, int cable, const tgTags& tags, const Config& config, double startLength);
    tgWorldB200Impl* m_world;
    int m_cable;
    Config m_config;
    double m_startLength;
    SpringCableActuatorHistory m_history;
};

class tgBasicActuator : public tgSpringCableActuator, public tgSubject<tgBasicActuator> {
public:
    typedef tgSpringCableActuator::Config Config;
    tgBasicActuator(tgWorldB200Impl* world, int cable, const tgTags& tags, const Config& config, double startLength);
    virtual void setup(tgWorld& world);
    virtual void teardown();
    virtual void step(double dt); /* dt<=0 -> std::invalid_argument */
    virtual void setControlInput(double input);            /* negative -> std::invalid_argument */
    virtual void setControlInput(double input, double dt); /* + moveMotors(dt) on the device, all envs */
    void moveMotors(double dt);
private:
    double m_preferredLength;
};

class tgKinematicActuator : public tgSpringCableActuator, public tgSubject<tgKinematicActuator> {
public:
    struct Config : public tgSpringCableActuator::Config {
        Config(double s = 1000.0, double d = 10.0, double p = 0.0, double rad = 1.0, double moFric = 0.0, double moInert = 1.0,
               bool back = false, bool h = false, double mf = 1000.0, double tVel = 100.0, double mnAL = 0.1, double mnRL = 0.1,
               double rot = 0);
        double radius;
        double motorFriction;
        double motorInertia;
        bool backdrivable;
        double maxOmega;
        double maxTorque;
    };
    tgKinematicActuator(tgWorldB200Impl* world, int cable, const tgTags& tags, const Config& config, double startLength);
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
    virtual void setControlInput(double input);
    virtual const double getVelocity() const;
    const Config& getKinematicConfig() const { return m_kconfig; }
private:
    Config m_kconfig;
    double m_desiredTorque;
};

/* ------------------------------------------------------------------------------------------------
 * View and simulation  (src/core/tgSimView.h, tgSimulation.h:45-127)
 * ---------------------------------------------------------------------------------------------- */
class tgSimulation;

/* Headless view: owns the step size and the run loop. */
class tgSimView {
public:
    tgSimView(tgWorld& world, double stepSize = 1.0 / 120.0, double renderRate = 1.0 / 60.0);
    virtual ~tgSimView();
    tgWorld& world() { return m_world; }
    virtual void setup();
    virtual void teardown();
    virtual void run();
    virtual void run(int steps);
    virtual void reset();
    void setStepSize(double stepSize);
    double getStepSize() const { return m_stepSize; }
    void setRenderRate(double renderRate);
    void bindToSimulation(tgSimulation& simulation);
    void releaseFromSimulation();
protected:
    tgWorld& m_world;
    tgSimulation* m_pSimulation;
    double m_stepSize, m_renderRate;
    bool m_initialized;
};

class tgDataManager;

class tgSimulation {
public:
    explicit tgSimulation(tgSimView& view);
    ~tgSimulation();
    void step(double dt) const;  /* dt<=0 -> std::invalid_argument */
    void run() const;
    void run(int steps) const;
    void addModel(tgModel* pModel); /* takes ownership; calls pModel->setup(world) */
    /* takes ownership; calls setup(); stepped after the models in every step (tgSimulation.cpp:104-116,213-216) */
    void addDataManager(tgDataManager* pDataManager);
    void reset();
    tgWorld& getWorld() const;
    /* True when any model-level or actuator-level host observer is attached (forces per-step host round trips). */
    bool hasHostObservers() const;
    bool hasHostStepCode() const;
    /* batched fast path: `steps` steps with no host observer in the loop */
    void runFused(double dt, int steps) const;
private:
    void teardown();
    tgSimView& m_view;
    std::vector<tgModel*> m_models;
    std::vector<tgDataManager*> m_dataManagers;
};

#endif /* TGB_HOST_H */
