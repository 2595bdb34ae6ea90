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
    const std::vector<tgSpringCableActuator*>& getAllActuators() const { return m_actuators; }
private:
    std::vector<tgSpringCableActuator*> m_actuators;
};

/* SUPERball as shipped: 6 rods, 25 cables incl. the duplicated (2,4)/(4,2) (src/examples/SUPERball/T6Model.cpp:54-212) */
class T6Model : public tgSubject<T6Model>, public tgModel {
public:
    T6Model();
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
    virtual void teardown();
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
    const std::vector<tgBasicActuator*>& getAllMuscles() const { return m_actuators; }
private:
    std::vector<tgBasicActuator*> m_actuators;
};

/* Motor test rig: a static rod, a hanging rod, one cable (src/examples/motorModel/tsTestRig.cpp:53-199) */
class tsTestRig : public tgSubject<tsTestRig>, public tgModel {
public:
    explicit tsTestRig(bool kinematic);
    virtual void setup(tgWorld& world);
    virtual void step(double dt);
    virtual void teardown();
    const std::vector<tgSpringCableActuator*>& getAllMuscles() const { return m_muscles; }
private:
    bool m_kinematic;
    std::vector<tgSpringCableActuator*> m_muscles;
};

/* EscapeController::transformActions (EscapeController.cpp:142-177) + applyActions' dcOffset = 1 (:179-194):
   maps 8x4 numbers in [0,1] to amplitude / angular frequency / phase change / dc offset. out = 33 doubles
   (32 + initial length), the TGB_CTRL_ESCAPE_SINE parameter row. */
void tgbEscapeTransformActions(const double* actions01, double initialLength, double* out33);

#endif
