/* tgb_examples.cpp -- table-driven example robots; see tgb_examples.h for the reference lines. */
#include "tgb_examples.h"

#include <cmath>

namespace {

struct IndexPair { int a, b; };

template <std::size_t N>
void addPairs(tgStructure& s, const IndexPair (&p)[N], const char* tag) {
    for (std::size_t i = 0; i < N; ++i) s.addPair(p[i].a, p[i].b, tag);
}

/* the twelve rod ends of a six-strut icosahedron tensegrity, struts pairwise parallel to X, Y, Z */
void addSixStrutNodes(tgStructure& s, double rodLength, double rodSpace) {
    const double h = rodLength / 2;
    const double n[12][3] = {
        {-rodSpace, -h, 0}, {-rodSpace, h, 0}, {rodSpace, -h, 0}, {rodSpace, h, 0},
        {0, -rodSpace, -h}, {0, -rodSpace, h}, {0, rodSpace, -h}, {0, rodSpace, h},
        {-h, 0, rodSpace},  {h, 0, rodSpace},  {-h, 0, -rodSpace}, {h, 0, -rodSpace},
    };
    for (int i = 0; i < 12; ++i) s.addNode(n[i][0], n[i][1], n[i][2]);
}
const IndexPair kSixStrutRods[6] = {{0, 1}, {2, 3}, {4, 5}, {6, 7}, {8, 9}, {10, 11}};

} // namespace

/* ---- 3-prism ---------------------------------------------------------------------------------------- */
PrismModel::PrismModel() {}
void PrismModel::setup(tgWorld& world) {
    const double density = 0.2, radius = 0.31, stiffness = 1000.0, damping = 10.0, pretension = 500.0;
    const double edge = 10.0, width = 10.0, height = 20.0;
    const tgRod::Config rodConfig(radius, density);
    const tgSpringCableActuator::Config muscleConfig(stiffness, damping, pretension);

    tgStructure s;
    for (int level = 0; level < 2; ++level) {
        const double y = level ? height : 0.0;
        s.addNode(-edge / 2.0, y, 0);
        s.addNode(edge / 2.0, y, 0);
        s.addNode(0, y, width);
    }
    const IndexPair rods[3] = {{0, 4}, {1, 5}, {2, 3}};
    const IndexPair muscles[9] = {{0, 1}, {1, 2}, {2, 0}, {3, 4}, {4, 5}, {5, 3}, {0, 3}, {1, 4}, {2, 5}};
    addPairs(s, rods, "rod");
    addPairs(s, muscles, "muscle");
    s.move(btVector3(0, 10, 0));

    tgBuildSpec spec;
    spec.addBuilder("rod", new tgRodInfo(rodConfig));
    spec.addBuilder("muscle", new tgBasicActuatorInfo(muscleConfig));
    tgStructureInfo structureInfo(s, spec);
    structureInfo.buildInto(*this, world);

    m_actuators = tgCast::filter<tgModel, tgSpringCableActuator>(getDescendants());
    notifySetup();
    tgModel::setup(world);
}
void PrismModel::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    notifyStep(dt);
    tgModel::step(dt);
}
void PrismModel::teardown() {
    notifyTeardown();
    m_actuators.clear();
    tgModel::teardown();
}

/* ---- SUPERball (T6) ---------------------------------------------------------------------------------- */
T6Model::T6Model() {}
void T6Model::setup(tgWorld& world) {
    const double density = 0.688, radius = 0.31, stiffness = 613.0, damping = 200.0, rodLength = 16.84, rodSpace = 7.5;
    const double friction = 0.99, rollFriction = 0.01, restitution = 0.0, pretension = 2452.0;
    const bool hist = false;
    const double maxTens = 100000, targetVelocity = 10000;
    const tgRod::Config rodConfig(radius, density, friction, rollFriction, restitution);
    tgBasicActuator::Config muscleConfig(stiffness, damping, pretension, hist, maxTens, targetVelocity);

    tgStructure s;
    addSixStrutNodes(s, rodLength, rodSpace);
    addPairs(s, kSixStrutRods, "rod");
    /* 25 pairs: (4,2) repeats (2,4); the shipped model has both (T6Model.cpp:143,153) */
    const IndexPair muscles[25] = {{0, 4}, {0, 5}, {0, 8},  {0, 10}, {1, 6},  {1, 7},  {1, 8},  {1, 10}, {2, 4},
                                   {2, 5}, {2, 9}, {2, 11}, {3, 7},  {3, 6},  {3, 9},  {3, 11}, {4, 2},  {4, 10},
                                   {4, 11}, {5, 8}, {5, 9},  {6, 10}, {6, 11}, {7, 8},  {7, 9}};
    addPairs(s, muscles, "muscle");
    s.move(btVector3(0, 10, 0));
    s.addRotation(btVector3(0, 0, 0), btVector3(0, 1, 0), M_PI / 2);

    tgBuildSpec spec;
    spec.addBuilder("rod", new tgRodInfo(rodConfig));
    spec.addBuilder("muscle", new tgBasicActuatorInfo(muscleConfig));
    tgStructureInfo structureInfo(s, spec);
    structureInfo.buildInto(*this, world);

    m_actuators = tgCast::filter<tgModel, tgBasicActuator>(getDescendants());
    notifySetup();
    tgModel::setup(world);
}
void T6Model::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    notifyStep(dt);
    tgModel::step(dt);
}
void T6Model::teardown() {
    notifyTeardown();
    m_actuators.clear();
    tgModel::teardown();
}

/* ---- craterEscape SUPERball --------------------------------------------------------------------------- */
EscapeModel::EscapeModel() {}
void EscapeModel::setup(tgWorld& world) {
    const double density = 0.40374, radius = 0.40, stiffness = 3152.36, damping = 50.0, rodLength = 17.00, rodSpace = 17.00 / 4;
    const double friction = 0.99, rollFriction = 0.01, restitution = 0.0;
    const bool hist = false;
    const double rotation = 0, maxTens = 100000, targetVelocity = 10000;
    const tgRod::Config rodConfig(radius, density, friction, rollFriction, restitution);
    /* Positional slip kept from the reference (EscapeModel.cpp:184-185): `hist` lands in the pretension slot and
       `rotation` in the hist slot, i.e. pretension 0, hist false. */
    tgBasicActuator::Config muscleConfig(stiffness, damping, hist, rotation, maxTens, targetVelocity);

    tgStructure s;
    addSixStrutNodes(s, rodLength, rodSpace);
    addPairs(s, kSixStrutRods, "rod");
    const IndexPair muscles[24] = {{0, 4}, {0, 5},  {0, 8},  {0, 10}, {1, 6},  {1, 7},  {1, 8}, {1, 10},
                                   {2, 4}, {2, 5},  {2, 9},  {2, 11}, {3, 7},  {3, 6},  {3, 9}, {3, 11},
                                   {4, 10}, {4, 11}, {5, 8},  {5, 9},  {6, 10}, {6, 11}, {7, 8}, {7, 9}};
    addPairs(s, muscles, "muscle");
    s.move(btVector3(0, 10, 0));
    s.addRotation(btVector3(0, 0, 0), btVector3(0, 1, 0), M_PI / 2);

    tgBuildSpec spec;
    spec.addBuilder("rod", new tgRodInfo(rodConfig));
    spec.addBuilder("muscle", new tgBasicActuatorInfo(muscleConfig));
    tgStructureInfo structureInfo(s, spec);
    structureInfo.buildInto(*this, world);

    m_actuators = tgCast::filter<tgModel, tgBasicActuator>(getDescendants());
    notifySetup();
    tgModel::setup(world);
}
void EscapeModel::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    notifyStep(dt);
    tgModel::step(dt);
}
void EscapeModel::teardown() {
    notifyTeardown();
    m_actuators.clear();
    tgModel::teardown();
}

/* ---- motor test rig ------------------------------------------------------------------------------------ */
tsTestRig::tsTestRig(bool kinematic) : m_kinematic(kinematic) {}
void tsTestRig::setup(tgWorld& world) {
    const double density = 2, radius = 0.31, stiffness = 1000.0, damping = 10.0;
    const tgRod::Config rodConfig(radius, density);
    const tgRod::Config rodConfig2(radius, 0.0); /* density 0 -> mass 0 -> static body */

    tgStructure s;
    s.addNode(0.0, -1.0, 0.0);
    s.addNode(0.0, 0.0, 0.0);
    s.addNode(0.0, 10.0, 0.0);
    s.addNode(0.0, 11.0, 0.0);
    s.addPair(0, 1, "rod");
    s.addPair(2, 3, "rod2");
    s.addPair(1, 2, "muscle");
    s.move(btVector3(0, 5, 0));

    tgBuildSpec spec;
    spec.addBuilder("rod", new tgRodInfo(rodConfig));
    spec.addBuilder("rod2", new tgRodInfo(rodConfig2));
    if (m_kinematic) spec.addBuilder("muscle", new tgKinematicActuatorInfo(tgKinematicActuator::Config(stiffness, damping)));
    else spec.addBuilder("muscle", new tgBasicActuatorInfo(tgBasicActuator::Config(stiffness, damping)));
    tgStructureInfo structureInfo(s, spec);
    structureInfo.buildInto(*this, world);

    m_muscles = tgCast::filter<tgModel, tgSpringCableActuator>(getDescendants());
    notifySetup();
    tgModel::setup(world);
}
void tsTestRig::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    notifyStep(dt);
    tgModel::step(dt);
}
void tsTestRig::teardown() {
    notifyTeardown();
    m_muscles.clear();
    tgModel::teardown();
}

/* ---- sine controller parameter map ---------------------------------------------------------------------- */
void tgbEscapeTransformActions(const double* a, double L0, double* out) {
    const double pretension = 0.90, maxStringLengthFactor = 0.50;
    const double mins[4] = {L0 * (pretension - maxStringLengthFactor), 0.3, -1 * M_PI, L0};
    const double maxes[4] = {L0 * (pretension + maxStringLengthFactor), 20, M_PI, L0};
    for (int c = 0; c < 8; ++c)
        for (int j = 0; j < 4; ++j) out[c * 4 + j] = a[c * 4 + j] * (maxes[j] - mins[j]) + mins[j];
    for (int c = 0; c < 8; ++c) out[c * 4 + 3] = 1; /* applyActions overwrites dcOffset with 1 */
    out[32] = L0;
}
