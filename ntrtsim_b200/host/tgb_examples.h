/*
 * tgb_examples.h -- the shipped example robots of the hot-path configs, written against the host API
 * of this repo (tgb_host.h / tgb_creator.h).  Constants come from the reference examples cited per class;
 * the models are table-driven here.
 */
#ifndef TGB_EXAMPLES_H
#define TGB_EXAMPLES_H

#include "tgb_creator.h"

/* 3 rods, 9 cables (src/examples/3_prism/PrismModel.cpp:50-168) */
class PrismModel : public tgSubject<PrismModel>, public tgModel {
public:
    PrismModel();
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
    virtual void teardown();
    virtual bool stepsOnHost() const { return numObservers() > 0; } /* model-level observers run on the host every step */
    const std::vector<tgSpringCableActuator*>& getAllActuators() const { return m_actuators; }
private:
    std::vector<tgSpringCableActuator*> m_actuators;
};

/* A rod and a box joined by one contact cable that runs into a third rod (src/examples/contactCables/ContactCableDemo.cpp:49-128; the
   model of test_integration/MuscleNP/MuscleNP_test.cpp).  The reference's setup also gives the first two bodies a central impulse of
   (0.4, 0, 0.4) through their btRigidBody; here the caller sets the corresponding velocities in the rod state. */
class ContactCableDemo : public tgSubject<ContactCableDemo>, public tgModel {
public:
    ContactCableDemo();
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
    virtual void teardown();
    const std::vector<tgSpringCableActuator*>& getAllMuscles() const { return m_muscles; }
private:
    std::vector<tgSpringCableActuator*> m_muscles;
};

/* SUPERball as shipped: 6 rods, 25 cables incl. the duplicated (2,4)/(4,2) (src/examples/SUPERball/T6Model.cpp:54-212) */
class T6Model : public tgSubject<T6Model>, public tgModel {
public:
    T6Model();
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
    virtual void teardown();
    virtual bool stepsOnHost() const { return numObservers() > 0; } /* model-level observers run on the host every step */
    const std::vector<tgBasicActuator*>& getAllActuators() const { return m_actuators; }
private:
    std::vector<tgBasicActuator*> m_actuators;
};

/* 24-cable SUPERball of craterEscape (src/examples/craterEscape/EscapeModel.cpp:55-214) */
class EscapeModel : public tgSubject<EscapeModel>, public tgModel {
public:
    EscapeModel();
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
    virtual void teardown();
    virtual bool stepsOnHost() const { return numObservers() > 0; } /* model-level observers run on the host every step */
    const std::vector<tgBasicActuator*>& getAllMuscles() const { return m_actuators; }
private:
    std::vector<tgBasicActuator*> m_actuators;
};

/* Crater obstacles (src/models/obstacles/tgCraterDeep.cpp:40-169, tgCraterShallow.cpp): four static tgBox walls (density 0) standing
   around `origin`, 20 from it, from 15 below to 20 above it; half width and half thickness 15. */
class tgCraterDeep : public tgModel {
public:
    tgCraterDeep();
    explicit tgCraterDeep(btVector3 center);
    virtual void setup(tgWorld& world);
    virtual void step(double dt); /* dt <= 0 -> std::invalid_argument */
protected:
    tgCraterDeep(btVector3 center, double halfSize);
private:
    btVector3 m_origin;
    double m_halfSize;
};
/* tgCraterShallow (src/models/obstacles/tgCraterShallow.cpp:72-168) is not a smaller tgCraterDeep: its four static boxes (half width
   and half thickness 10) run along the DIAGONAL from (-25, -7, -25) to (25, 7, 25), turned an eighth about the vertical through
   (25, 7, 25); the structure so far is turned a quarter about the vertical through `origin` after each, and all sink 5 into the ground --
   low ramps leaning outwards instead of walls.  (tests/test_reference_sources.py holds this against the reference's own source.) */
class tgCraterShallow : public tgModel {
public:
    tgCraterShallow();
    explicit tgCraterShallow(btVector3 center);
    virtual void setup(tgWorld& world);
    virtual void step(double dt); /* dt <= 0 -> std::invalid_argument */
private:
    btVector3 m_origin;
};

/* A flight of stairs (src/models/obstacles/tgStairs.{h,cpp}): nStairs static boxes, each `stepWidth` further and `stepHeight`
   higher than the one before, turned by `angle` about the vertical and moved to `origin`. */
class tgStairs : public tgModel {
public:
    struct Config {
        Config(btVector3 origin = btVector3(0.0, 0.0, 0.0), btScalar friction = 0.5, btScalar restitution = 0.0, std::size_t nStairs = 10,
               double stairWidth = 20.0, double stepWidth = 5.0, double stepHeight = 1.0, double angle = 0.0);
        btVector3 m_origin;
        btScalar m_friction, m_restitution;
        std::size_t m_nBlocks;
        double m_length, m_width, m_height, m_angle;
    };
    tgStairs();
    explicit tgStairs(const Config& config);
    virtual void setup(tgWorld& world);
    virtual void step(double dt); /* dt <= 0 -> std::invalid_argument */
private:
    Config m_config;
};

/* A field of static blocks at random places (src/models/obstacles/tgBlockField.{h,cpp}; positions from rand() like there). An
   environment holds at most 32 bodies, so the reference's default of 500 blocks does not fit: pass a smaller nBlocks. */
class tgBlockField : public tgModel {
public:
    struct Config {
        Config(btVector3 origin = btVector3(0.0, 0.0, 0.0), btScalar friction = 0.5, btScalar restitution = 0.0,
               btVector3 minPos = btVector3(-100.0, 0.0, -100.0), btVector3 maxPos = btVector3(100.0, 0.0, 100.0), std::size_t nBlocks = 500,
               double blockLength = 5.0, double blockWidth = 5.0, double blockHeight = 5.0);
        btVector3 m_origin;
        btScalar m_friction, m_restitution;
        btVector3 m_minPos, m_maxPos;
        std::size_t m_nBlocks;
        double m_length, m_width, m_height;
    };
    tgBlockField();
    explicit tgBlockField(const Config& config);
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
private:
    Config m_config;
};

/* One long static box through `origin` (src/models/obstacles/tgWall.{h,cpp}: class Wall): 200 long, half extents 5 x 1. */
class Wall : public tgModel {
public:
    Wall();
    explicit Wall(btVector3 origin);
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
private:
    btVector3 m_origin;
};

/* Motor test rig: a static rod, a hanging rod, one cable (src/examples/motorModel/tsTestRig.cpp:53-199) */
class tsTestRig : public tgSubject<tsTestRig>, public tgModel {
public:
    explicit tsTestRig(bool kinematic);
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
    virtual void teardown();
    virtual bool stepsOnHost() const { return true; } /* step() commands the motor every step (tsTestRig.cpp:181-188) */
    const std::vector<tgSpringCableActuator*>& getAllMuscles() const { return m_muscles; }
private:
    bool m_kinematic;
    std::vector<tgSpringCableActuator*> m_muscles;
};

/* 3-segment tetrahedral spine of the learning examples: every tetrahedron's 6 rods auto-compound into one body, 6 cables
   between consecutive segments (src/examples/learningSpines/TetraSpine/TetraSpineLearningModel.cpp:54-213) */
class TetraSpineLearningModel : public tgSubject<TetraSpineLearningModel>, public tgModel {
public:
    explicit TetraSpineLearningModel(std::size_t segments);
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
    virtual void teardown();
    virtual bool stepsOnHost() const { return numObservers() > 0; } /* model-level observers run on the host every step */
    const std::vector<tgSpringCableActuator*>& getAllMuscles() const { return m_allMuscles; }
    std::size_t getSegments() const { return m_segments; }
    const std::vector<tgModel*>& getAllSegments() const { return m_allSegments; }
private:
    std::size_t m_segments;
    std::vector<tgSpringCableActuator*> m_allMuscles;
    std::vector<tgModel*> m_allSegments;
};

/* 12-segment tetrahedral-complex ("Flemons") spine on kinematic actuators: every vertebra's 4 rods (centre to the corners of a
   tetrahedron) auto-compound into one body, 8 tgKinematicActuator cables between consecutive vertebrae
   (src/examples/learningSpines/TetrahedralComplex/FlemonsSpineModelLearning.cpp:51-206, USE_KINEMATIC) */
class FlemonsSpineModelLearning : public tgSubject<FlemonsSpineModelLearning>, public tgModel {
public:
    explicit FlemonsSpineModelLearning(std::size_t segments);
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
    virtual void teardown();
    virtual bool stepsOnHost() const { return numObservers() > 0; }
    const std::vector<tgSpringCableActuator*>& getAllMuscles() const { return m_allMuscles; }
    std::size_t getSegments() const { return m_segments; }
    const std::vector<tgModel*>& getAllSegments() const { return m_allSegments; }
private:
    std::size_t m_segments;
    std::vector<tgSpringCableActuator*> m_allMuscles;
    std::vector<tgModel*> m_allSegments;
};

/* Hardware-faithful 3-segment tetrahedral spine of the ICRA 2015 hardware comparison: every segment is a compound of 8 rods and 6
   point masses (spheres at the tagged nodes: corners, tip, PCBs), 6 tgBasicActuator cables between consecutive segments with
   pretension 2000 (top) / 700 (left, right) (src/examples/IROS_2015/TetraSpineStatic/TetraSpineStaticModel_hf.cpp:51-330,
   test_integration/ICRA2015Tests/ICRA2015_test.cpp:77-101) */
class TetraSpineStaticModel_hf : public tgSubject<TetraSpineStaticModel_hf>, public tgModel {
public:
    explicit TetraSpineStaticModel_hf(std::size_t segments);
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
    virtual void teardown();
    virtual bool stepsOnHost() const { return numObservers() > 0; }
    const std::vector<tgSpringCableActuator*>& getAllMuscles() const { return m_allMuscles; }
    std::size_t getSegments() const { return m_segments; }
    const std::vector<tgModel*>& getAllSegments() const { return m_allSegments; }
private:
    std::size_t m_segments;
    std::vector<tgSpringCableActuator*> m_allMuscles;
    std::vector<tgModel*> m_allSegments;
};

/* ---- CPG + impedance control of the spine examples, flattened for tgb_set_cpg_controller ------------------------- */
/* BaseSpineCPGControl::Config (BaseSpineCPGControl.h:60-77) with its defaults */
struct tgbSpineControlConfig {
    int segmentSpan, theirMuscles, ourMuscles, params, segmentNumber;
    double controlTime, lowAmp, highAmp, lowPhase, highPhase, lowFreq, highFreq;
    tgbSpineControlConfig(int ss, int tm, int om, int param, int segnum = 6, double ct = 0.1);
    int tablePairs() const { return segmentSpan * theirMuscles * ourMuscles; }
};
/* tgCPGActuatorControl::setConnectivity (tgCPGActuatorControl.cpp:121-173): couplings of every cable's CPG node in CSR form;
   pair[] indexes the [segmentSpan][theirMuscles][ourMuscles] edge table. */
void tgbSpineCpgTopology(const std::vector<tgb_cable_desc>& cables, const tgbSpineControlConfig& cfg, std::vector<int32_t>& begin,
                         std::vector<int32_t>& target, std::vector<int32_t>& pair);
/* BaseSpineCPGControl::scaleEdgeActions (BaseSpineCPGControl.cpp:363-434): the action list (popped from the back) fills the upper
   triangle of the edge table. actions: [n][2] in [0,1]; table: [tablePairs()][2] (weight, phase); returns the number consumed. */
int tgbSpineScaleEdgeActions(const double* actions, int n_actions, const tgbSpineControlConfig& cfg, double* table);
/* BaseSpineCPGControl::scaleNodeActions (BaseSpineCPGControl.cpp:435-468): [frequency, amplitude] of controller 0 */
void tgbSpineScaleNodeActions(const double* actions2, const tgbSpineControlConfig& cfg, double* out2);

/* EscapeController::transformActions (EscapeController.cpp:142-177) + applyActions' dcOffset = 1 (:179-194):
   maps 8x4 numbers in [0,1] to amplitude / angular frequency / phase change / dc offset. out = 33 doubles
   (32 + initial length), the TGB_CTRL_ESCAPE_SINE parameter row. */
void tgbEscapeTransformActions(const double* actions01, double initialLength, double* out33);

#endif
