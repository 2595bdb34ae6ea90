/* tgb_examples.cpp -- table-driven example robots; see tgb_examples.h for the reference lines. */
#include "tgb_examples.h"

#include <cmath>
#include <sstream>

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

/* ---- contact cable demo ------------------------------------------------------------------------------ */
ContactCableDemo::ContactCableDemo() : tgModel() {}
void ContactCableDemo::setup(tgWorld& world) {
    const double rodDensity = 1, rodRadius = 0.25;
    const tgRod::Config rodConfig(rodRadius, rodDensity);
    const tgBox::Config boxConfig(rodRadius, rodRadius, rodDensity);

    tgStructure s;
    s.addNode(0, 2, 0);
    s.addNode(0, 4, 0);
    s.addNode(0, 12, 0);
    s.addNode(0, 14, 0);
    /* the rod the cable runs into */
    s.addNode(-1, 8, 5);
    s.addNode(1, 8, 5);
    s.addPair(0, 1, "rod");
    s.addPair(2, 3, "box");
    s.addPair(4, 5, "rod");
    s.addPair(1, 2, "muscle");
    s.addRotation(btVector3(0.0, 0.0, 0.0), btVector3(0.0, 1.0, 0.0), M_PI / 4);
    tgSpringCableActuator::Config muscleConfig(1000, 0, 0.0, false, 600000000);

    tgBuildSpec spec;
    spec.addBuilder("box", new tgBoxInfo(boxConfig));
    spec.addBuilder("rod", new tgRodInfo(rodConfig));
    spec.addBuilder("muscle", new tgBasicContactCableInfo(muscleConfig));
    tgStructureInfo structureInfo(s, spec);
    structureInfo.buildInto(*this, world);

    m_muscles = tgCast::filter<tgModel, tgSpringCableActuator>(getDescendants());
    notifySetup();
    tgModel::setup(world);
}
void ContactCableDemo::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    notifyStep(dt);
    tgModel::step(dt);
}
void ContactCableDemo::teardown() {
    notifyTeardown();
    m_muscles.clear();
    tgModel::teardown();
}

/* ---- crater obstacles -------------------------------------------------------------------------------- */
tgCraterDeep::tgCraterDeep() : m_origin(0, 0, 0), m_halfSize(15.0) {}
tgCraterDeep::tgCraterDeep(btVector3 center) : m_origin(center), m_halfSize(15.0) {}
tgCraterDeep::tgCraterDeep(btVector3 center, double halfSize) : m_origin(center), m_halfSize(halfSize) {}
void tgCraterDeep::setup(tgWorld& world) {
    /* friction 1, rolling friction 0.01, restitution 0.2 like the rods they meet; density 0 makes the boxes static */
    const tgBox::Config wallConfig(m_halfSize, m_halfSize, 0.0, 1.0, 0.01, 0.2);
    tgStructure s;
    for (int wall = 0; wall < 4; ++wall) {
        /* every wall is entered at x = 20 and the structure so far is turned a quarter about the vertical through the origin
           before the pair is made, so the four walls end up a quarter turn apart (tgCraterDeep.cpp:124-145) */
        s.addNode(20.0, -10.0, 0.0, "node");
        s.addNode(20.0, 25.0, 0.0, "node");
        s.addRotation(m_origin, btVector3(0, 1, 0), M_PI / 2);
        s.addPair(2 * wall, 2 * wall + 1, "box");
    }
    s.move(btVector3(0, -5, 0)); /* sunk into the ground */
    tgBuildSpec spec;
    spec.addBuilder("box", new tgBoxInfo(wallConfig));
    tgStructureInfo structureInfo(s, spec);
    structureInfo.buildInto(*this, world);
    tgModel::setup(world);
}
void tgCraterDeep::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    tgModel::step(dt);
}

tgCraterShallow::tgCraterShallow() : m_origin(0, 0, 0) {}
tgCraterShallow::tgCraterShallow(btVector3 center) : m_origin(center) {}
void tgCraterShallow::setup(tgWorld& world) {
    const tgBox::Config rampConfig(10.0, 10.0, 0.0, 1.0, 0.01, 0.2); /* density 0: static */
    /* one ramp: the two ends of a box diagonal, 20 + half a width out in x and z and half a height + 2 up and down, both turned an
       eighth about the vertical through the point half way between them measured from zero (tgCraterShallow.cpp:141-167) */
    const double reach = 20.0 + 10.0 / 2.0, rise = 10.0 / 2.0 + 2.0;
    const btVector3 pivot((reach + reach) / 2, (rise + rise) / 2, (reach + reach) / 2);
    tgNode low(-reach, -rise, -reach, "node"), high(reach, rise, reach, "node");
    low.addRotation(pivot, btVector3(0, 1, 0), M_PI / 4);
    high.addRotation(pivot, btVector3(0, 1, 0), M_PI / 4);
    tgStructure s;
    for (int ramp = 0; ramp < 4; ++ramp) {
        s.addNode(low);
        s.addNode(high);
        s.addRotation(m_origin, btVector3(0, 1, 0), M_PI / 2); /* everything entered so far turns a quarter */
        s.addPair(2 * ramp, 2 * ramp + 1, "box");
    }
    s.move(btVector3(0, -5, 0)); /* sunk into the ground */
    tgBuildSpec spec;
    spec.addBuilder("box", new tgBoxInfo(rampConfig));
    tgStructureInfo structureInfo(s, spec);
    structureInfo.buildInto(*this, world);
    tgModel::setup(world);
}
void tgCraterShallow::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    tgModel::step(dt);
}

namespace {
/* the common tail of the obstacle models: every pair tagged "box" becomes a static tgBox */
void buildStaticBoxes(tgModel& self, tgStructure& s, const tgBox::Config& boxConfig, tgWorld& world) {
    tgBuildSpec spec;
    spec.addBuilder("box", new tgBoxInfo(boxConfig));
    tgStructureInfo structureInfo(s, spec);
    structureInfo.buildInto(self, world);
}
} // namespace

tgStairs::Config::Config(btVector3 origin, btScalar friction, btScalar restitution, std::size_t nStairs, double stairWidth, double stepWidth,
                         double stepHeight, double angle)
    : m_origin(origin), m_friction(friction), m_restitution(restitution), m_nBlocks(nStairs), m_length(stairWidth), m_width(stepWidth),
      m_height(stepHeight), m_angle(angle) {}
tgStairs::tgStairs() {}
tgStairs::tgStairs(const Config& config) : m_config(config) {}
void tgStairs::setup(tgWorld& world) {
    /* density and rolling friction zero; the Config's full sizes halved into tgBox half extents (tgStairs.cpp:82-84) */
    const tgBox::Config boxConfig(m_config.m_width / 2.0, m_config.m_height / 2.0, 0.0, m_config.m_friction, 0.0, m_config.m_restitution);
    tgStructure s;
    tgNode from(0.0, 0.0, 0.0), to(0.0, 0.0, m_config.m_length);
    const btVector3 rise(m_config.m_width, m_config.m_height, 0.0);
    for (std::size_t i = 0; i < m_config.m_nBlocks; ++i) {
        s.addNode(from);
        s.addNode(to);
        s.addPair((int)(2 * i), (int)(2 * i + 1), "box");
        from += rise;
        to += rise;
    }
    s.addRotation(btVector3(0.0, 0.0, 0.0), btVector3(0.0, 1.0, 0.0), m_config.m_angle);
    s.move(m_config.m_origin);
    buildStaticBoxes(*this, s, boxConfig, world);
    tgModel::setup(world);
}
void tgStairs::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    tgModel::step(dt);
}

tgBlockField::Config::Config(btVector3 origin, btScalar friction, btScalar restitution, btVector3 minPos, btVector3 maxPos, std::size_t nBlocks,
                             double blockLength, double blockWidth, double blockHeight)
    : m_origin(origin), m_friction(friction), m_restitution(restitution), m_minPos(minPos), m_maxPos(maxPos), m_nBlocks(nBlocks),
      m_length(blockLength), m_width(blockWidth), m_height(blockHeight) {}
tgBlockField::tgBlockField() {}
tgBlockField::tgBlockField(const Config& config) : m_config(config) {}
void tgBlockField::setup(tgWorld& world) {
    const tgBox::Config boxConfig(m_config.m_width / 2.0, m_config.m_height / 2.0, 0.0, m_config.m_friction, 0.0, m_config.m_restitution);
    tgStructure s;
    const btVector3 fieldSize = m_config.m_maxPos - m_config.m_minPos;
    for (std::size_t i = 0; i < m_config.m_nBlocks; ++i) {
        /* three draws per block, x then y then z (tgBlockField.cpp:131-145) */
        const double xOffset = fieldSize.getX() * rand() / RAND_MAX;
        const double yOffset = fieldSize.getY() * rand() / RAND_MAX;
        const double zOffset = fieldSize.getZ() * rand() / RAND_MAX;
        tgNode position(m_config.m_minPos + btVector3(xOffset, yOffset, zOffset));
        s.addNode(position);
        position += btVector3(0.0, 0.0, m_config.m_length);
        s.addNode(position);
        s.addPair((int)(2 * i), (int)(2 * i + 1), "box");
    }
    s.move(m_config.m_origin);
    buildStaticBoxes(*this, s, boxConfig, world);
    tgModel::setup(world);
}
void tgBlockField::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    tgModel::step(dt);
}

Wall::Wall() : m_origin(0, 0, 0) {}
Wall::Wall(btVector3 origin) : m_origin(origin) {}
void Wall::setup(tgWorld& world) {
    const tgBox::Config boxConfig(5.0, 1.0, 0.0, 1.0, 0.01, 0.2);
    tgStructure s;
    s.addNode(-100.0, 0.0, 0.0, "node");
    s.addNode(100.0, 0.0, 0.0, "node");
    s.addPair(0, 1, "box");
    s.move(m_origin);
    buildStaticBoxes(*this, s, boxConfig, world);
    tgModel::setup(world);
}
void Wall::step(double dt) {
    if (dt <= 0.0) throw std::invalid_argument("dt is not positive");
    tgModel::step(dt);
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
    /* tsTestRig.cpp:181-188 */
    if (m_kinematic) m_muscles[0]->setControlInput(-580.0);
    else m_muscles[0]->setControlInput(5.0, dt);
    tgModel::step(dt);
}
void tsTestRig::teardown() {
    notifyTeardown();
    m_muscles.clear();
    tgModel::teardown();
}

/* ---- tetrahedral spine -------------------------------------------------------------------------------------------- */
namespace {
double round5(double d) {
    /* tgUtil::round(d, 5), tgUtil.h:354-359 */
    const int m = static_cast<int>(std::pow(10.0, 5));
    return std::floor(d * m + 0.5) / m;
}
std::string numbered(const std::string& s, int i) {
    std::ostringstream ss;
    ss << s << i;
    return ss.str();
}
} // namespace

TetraSpineLearningModel::TetraSpineLearningModel(std::size_t segments) : m_segments(segments) {}
void TetraSpineLearningModel::setup(tgWorld& world) {
    const double edge = 38.1;
    const double height = round5(std::sqrt(3.0) / 2 * edge);
    tgStructure tetra;
    tetra.addNode(-edge / 2.0, 0, 0);                                         /* right */
    tetra.addNode(edge / 2.0, 0, 0);                                          /* left */
    tetra.addNode(0, height, 0);                                              /* top */
    tetra.addNode(0, height / 2.0, round5(std::sqrt(3.0) / 2.0 * height));    /* front */
    const IndexPair rods[6] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
    const char* rodTags[6] = {"back bottom rod", "back right rod", "front right rod", "back left rod", "front left rod", "front top rod"};
    for (int i = 0; i < 6; ++i) tetra.addPair(rods[i].a, rods[i].b, rodTags[i]);
    tetra.move(btVector3(0.0, 2.0, 100.0));

    tgStructure snake;
    const btVector3 offset(0, 0, -edge * 0.75);
    for (std::size_t i = 0; i < m_segments; ++i) {
        tgStructure* const t = new tgStructure(tetra);
        t->addTags(numbered("segment num", (int)i + 1));
        t->move((btScalar)(i + 1) * offset);
        snake.addChild(t);
    }
    const std::vector<tgStructure*>& children = snake.getChildren();
    for (std::size_t i = 1; i < children.size(); ++i) {
        const tgNodes& n0 = children[i - 1]->getNodes();
        const tgNodes& n1 = children[i]->getNodes();
        snake.addPair(n0[0], n1[0], "outer right muscle");
        snake.addPair(n0[1], n1[1], "outer left muscle");
        snake.addPair(n0[2], n1[2], "outer top muscle");
        snake.addPair(n0[0], n1[3], "inner right muscle");
        snake.addPair(n0[1], n1[3], "inner left muscle");
        snake.addPair(n0[2], n1[3], "inner top muscle");
    }

    const double density = .00311, radius = 0.635, friction = 0.5;
    const tgRod::Config rodConfig(radius, density, friction);
    tgBuildSpec spec;
    spec.addBuilder("rod", new tgRodInfo(rodConfig));
    /* positional arguments as shipped (TetraSpineLearningModel.cpp:192): pretension = false, hist = 0 */
    tgSpringCableActuator::Config muscleConfig(10000, 10, false, 0, 7000, 7.0);
    spec.addBuilder("muscle", new tgBasicActuatorInfo(muscleConfig));
    tgStructureInfo structureInfo(snake, spec);
    structureInfo.buildInto(*this, world);

    m_allMuscles = find<tgSpringCableActuator>(std::string("muscle"));
    m_allSegments = find<tgModel>(std::string("segment"));
    notifySetup();
    tgModel::setup(world);
}
void TetraSpineLearningModel::step(double dt) {
    notifyStep(dt);
    tgModel::step(dt);
}
void TetraSpineLearningModel::teardown() {
    notifyTeardown();
    tgModel::teardown();
    m_allMuscles.clear();
    m_allSegments.clear();
}

/* ---- hardware-faithful tetrahedral spine (rods + point masses) ----------------------------------------------------- */
TetraSpineStaticModel_hf::TetraSpineStaticModel_hf(std::size_t segments) : m_segments(segments) {}
void TetraSpineStaticModel_hf::setup(tgWorld& world) {
    const double edge = 38;
    const double height = round5(std::sqrt(3.0) / 2 * edge);
    tgStructure tetra;
    /* addNodes, TetraSpineStaticModel_hf.cpp:62-90 */
    tetra.addNode(-edge / 2.0, 0, 0, "base");                               /* right */
    tetra.addNode(edge / 2.0, 0, 0, "base");                                /* left */
    tetra.addNode(0, height, 0, "base");                                    /* top */
    tetra.addNode(0, edge / 2.0 * std::tan(M_PI / 6.0), 25.0, "tip");       /* front */
    {
        const tgNodes& old = tetra.getNodes();
        tgNode nn1((old[0] + old[2]) / 2.0);                                /* 4 mid */
        nn1.addTags(std::string("PCB"));
        tgNode nn2((old[1] + old[2]) / 2.0);                                /* 5 mid */
        nn2.addTags(std::string("PCB"));
        tetra.addNode(nn1);
        tetra.addNode(nn2);
    }
    /* addPairs, :92-102 */
    tetra.addPair(0, 1, "back bottom rod");
    tetra.addPair(0, 4, "back rightBottom rod");
    tetra.addPair(4, 2, "back rightTop rod");
    tetra.addPair(0, 3, "front right rod");
    tetra.addPair(1, 5, "back leftBottom rod");
    tetra.addPair(5, 2, "back leftTop rod");
    tetra.addPair(1, 3, "front left rod");
    tetra.addPair(2, 3, "front top rod");
    tetra.move(btVector3(0.0, 2.0, 100.0));

    tgStructure snake;
    const btVector3 offset(0, 0, -23.0);
    for (std::size_t i = 0; i < m_segments; ++i) {
        tgStructure* const t = new tgStructure(tetra);
        t->addTags(numbered("segment num", (int)i)); /* num0, num1, num2 select the point-mass builders below */
        t->move((btScalar)(i + 1) * offset);
        snake.addChild(t);
    }
    const std::vector<tgStructure*>& children = snake.getChildren();
    for (std::size_t i = 1; i < children.size(); ++i) {
        const tgNodes& n0 = children[i - 1]->getNodes();
        const tgNodes& n1 = children[i]->getNodes();
        snake.addPair(n0[0], n1[0], numbered("outer right muscle seg", (int)i));
        snake.addPair(n0[1], n1[1], numbered("outer left muscle seg", (int)i));
        snake.addPair(n0[2], n1[2], numbered("outer top muscle seg", (int)i));
        snake.addPair(n0[0], n1[3], numbered("inner right muscle seg", (int)i));
        snake.addPair(n0[1], n1[3], numbered("inner left muscle seg", (int)i));
        snake.addPair(n0[2], n1[3], numbered("inner top muscle seg", (int)i));
    }

    /* :262-313 */
    const double radius = 0.635 / 2.0;
    const double density = 0.0201 / (std::pow(radius, 2) * M_PI * edge);
    const double friction = 0.15;
    const tgRod::Config rodConfig(radius, density, friction);
    tgBuildSpec spec;
    spec.addBuilder("rod", new tgRodInfo(rodConfig));
    const double sphereVolume = 1000.0 * 4.0 / 3.0 * M_PI * std::pow(radius, 3); /* 1000 so that the numbers below are grams */
    struct { const char* tags; double grams; } masses[9] = {{"num0 base", 140.0}, {"num1 base", 180.0}, {"num2 base", 100.0},
                                                            {"num0 tip", 40.0},   {"num1 tip", 120.0},  {"num2 tip", 120.0},
                                                            {"PCB num0", 70.0},   {"PCB num1", 70.0},   {"PCB num2", 95.0}};
    for (int i = 0; i < 9; ++i) spec.addBuilder(masses[i].tags, new tgSphereInfo(tgSphere::Config(radius, masses[i].grams / sphereVolume, friction)));
    tgSpringCableActuator::Config muscleConfig(229.16, 20, 2000.0, true, 5000, 7.0, 1.0, 1.0);
    tgSpringCableActuator::Config muscleConfig2(229.16, 20, 700.0, true, 5000, 7.0, 1.0, 1.0);
    spec.addBuilder("top muscle", new tgBasicActuatorInfo(muscleConfig));
    spec.addBuilder("left muscle", new tgBasicActuatorInfo(muscleConfig2));
    spec.addBuilder("right muscle", new tgBasicActuatorInfo(muscleConfig2));
    tgStructureInfo structureInfo(snake, spec);
    structureInfo.buildInto(*this, world);

    m_allMuscles = find<tgSpringCableActuator>(std::string("muscle"));
    m_allSegments = find<tgModel>(std::string("segment"));
    notifySetup();
    tgModel::setup(world);
}
void TetraSpineStaticModel_hf::step(double dt) {
    if (dt < 0.0) throw std::invalid_argument("dt is not positive");
    notifyStep(dt);
    tgModel::step(dt);
}
void TetraSpineStaticModel_hf::teardown() {
    notifyTeardown();
    tgModel::teardown();
    m_allMuscles.clear();
    m_allSegments.clear();
}

/* ---- tetrahedral complex spine, kinematic actuators ----------------------------------------------------------------- */
FlemonsSpineModelLearning::FlemonsSpineModelLearning(std::size_t segments) : m_segments(segments) {}
void FlemonsSpineModelLearning::setup(tgWorld& world) {
    /* FlemonsSpineModelLearning.cpp:62-92 */
    const double density = 4.2 / 300.0, radius = 0.5, friction = 0.5, rollFriction = 0.0, restitution = 0.0;
    const tgRod::Config rodConfig(radius, density, friction, rollFriction, restitution);
    const double elasticity = 1000.0, damping = 10.0, pretension = 0.0, maxTens = 7000.0, maxSpeed = 12.0;
    const double mRad = 1.0, motorFriction = 10.0, motorInertia = 1.0;
    const bool backDrivable = false, history = false;
    tgKinematicActuator::Config motorConfig(elasticity, damping, pretension, mRad, motorFriction, motorInertia, backDrivable, history, maxTens,
                                            maxSpeed);
    const double v_size = 10.0;
    tgStructure tetra;
    tetra.addNode(0.0, 0.0, 0.0);             /* center */
    tetra.addNode(v_size, v_size, v_size);    /* front */
    tetra.addNode(v_size, -v_size, -v_size);  /* right */
    tetra.addNode(-v_size, v_size, -v_size);  /* back */
    tetra.addNode(-v_size, -v_size, v_size);  /* left */
    const char* rodTags[4] = {"front rod", "right rod", "back rod", "left rod"};
    for (int i = 0; i < 4; ++i) tetra.addPair(0, i + 1, rodTags[i]);
    tetra.move(btVector3(0.0, 15.0, 100.0));

    tgStructure snake;
    const btVector3 offset(0.0, 0.0, -v_size * 1.15);
    for (std::size_t i = 0; i < m_segments; ++i) {
        tgStructure* const p = new tgStructure(tetra);
        p->addTags(numbered("segment num", (int)i + 1));
        p->move((btScalar)(i + 1.0) * offset);
        snake.addChild(p);
    }
    const std::vector<tgStructure*>& children = snake.getChildren();
    for (std::size_t i = 1; i < children.size(); ++i) {
        const tgNodes& n0 = children[i - 1]->getNodes();
        const tgNodes& n1 = children[i]->getNodes();
        const std::string seg = numbered(" seg", (int)i - 1) + numbered(" seg", (int)i);
        snake.addPair(n0[1], n1[1], "outer front muscle" + seg);
        snake.addPair(n0[2], n1[2], "outer right muscle" + seg);
        snake.addPair(n0[3], n1[3], "outer back muscle" + seg);
        snake.addPair(n0[4], n1[4], "outer top muscle" + seg);
        snake.addPair(n0[2], n1[1], "inner front muscle" + seg);
        snake.addPair(n0[2], n1[4], "inner right muscle" + seg);
        snake.addPair(n0[3], n1[1], "inner left muscle" + seg);
        snake.addPair(n0[3], n1[4], "inner back muscle" + seg);
    }
    tgBuildSpec spec;
    spec.addBuilder("rod", new tgRodInfo(rodConfig));
    spec.addBuilder("muscle", new tgKinematicActuatorInfo(motorConfig));
    tgStructureInfo structureInfo(snake, spec);
    structureInfo.buildInto(*this, world);

    m_allMuscles = tgCast::filter<tgModel, tgSpringCableActuator>(getDescendants());
    m_allSegments = find<tgModel>(std::string("segment"));
    notifySetup();
    tgModel::setup(world);
}
void FlemonsSpineModelLearning::step(double dt) {
    notifyStep(dt);
    tgModel::step(dt);
}
void FlemonsSpineModelLearning::teardown() {
    notifyTeardown();
    tgModel::teardown();
    m_allMuscles.clear();
    m_allSegments.clear();
}

tgbSpineControlConfig::tgbSpineControlConfig(int ss, int tm, int om, int param, int segnum, double ct)
    : segmentSpan(ss), theirMuscles(tm), ourMuscles(om), params(param), segmentNumber(segnum), controlTime(ct), lowAmp(0), highAmp(30),
      lowPhase(-1 * M_PI), highPhase(M_PI), lowFreq(0.0), highFreq(30.0) {
    if (ss <= 0) throw std::invalid_argument("segmentSpan parameter is negative.");
    else if (tm <= 0) throw std::invalid_argument("theirMuscles parameter is negative.");
    else if (om <= 0) throw std::invalid_argument("Our Muscles parameter is negative.");
    else if (param <= 0) throw std::invalid_argument("Edge parameters is negative.");
    else if (segnum < 0) throw std::invalid_argument("Segment number is negative.");
    else if (ct < 0.0) throw std::invalid_argument("control time is negative.");
}

void tgbSpineCpgTopology(const std::vector<tgb_cable_desc>& cables, const tgbSpineControlConfig& cfg, std::vector<int32_t>& begin,
                         std::vector<int32_t>& target, std::vector<int32_t>& pair) {
    const int n = (int)cables.size();
    const int muscleSize = cfg.theirMuscles; /* edgeParams.shape()[1] */
    begin.assign(1, 0);
    target.clear();
    pair.clear();
    for (int m = 0; m < n; ++m) {
        for (int t = 0; t < n; ++t) {
            if (t == m) continue;
            /* "all to all" connectivity for shared rigid bodies */
            if (cables[m].body_a == cables[t].body_a || cables[m].body_b == cables[t].body_b || cables[m].body_a == cables[t].body_b ||
                cables[m].body_b == cables[t].body_a) {
                const int rp = ((m - t) / muscleSize) + 1; /* integer division: -1 behind, 0 same row, 1 ahead */
                const int j = m % muscleSize, k = t % muscleSize;
                if (rp < 0 || rp >= cfg.segmentSpan) throw std::runtime_error("CPG coupling outside the segment span");
                const int lo = j > k ? k : j, hi = j > k ? j : k;
                target.push_back(t);
                pair.push_back((rp * cfg.theirMuscles + lo) * cfg.ourMuscles + hi);
            }
        }
        begin.push_back((int32_t)target.size());
    }
}

int tgbSpineScaleEdgeActions(const double* actions, int n_actions, const tgbSpineControlConfig& cfg, double* table) {
    const double range = cfg.highPhase - cfg.lowPhase;
    for (int i = 0; i < cfg.tablePairs() * 2; ++i) table[i] = 0.0;
    int remaining = n_actions;
    int i = 0, j = 0, k = 0;
    while (i < cfg.segmentSpan) {
        while (j < cfg.theirMuscles) {
            while (k < cfg.ourMuscles) {
                if (remaining == 0) break; /* "ran out before table populated!" */
                if (!(i == 1 && j == k)) {
                    const double* edgeParam = actions + 2 * (remaining - 1); /* actions.back() */
                    double* e = table + 2 * ((i * cfg.theirMuscles + j) * cfg.ourMuscles + k);
                    e[0] = edgeParam[0];
                    e[1] = edgeParam[1] * (range) + cfg.lowPhase;
                    --remaining;
                }
                k++;
            }
            j++;
            k = j;
        }
        j = 0;
        k = 0;
        i++;
    }
    return n_actions - remaining;
}

void tgbSpineScaleNodeActions(const double* a, const tgbSpineControlConfig& cfg, double* out) {
    out[0] = (a[0] * (cfg.highFreq - cfg.lowFreq)) + cfg.lowFreq;
    out[1] = (a[1] * (cfg.highAmp - cfg.lowAmp)) + cfg.lowAmp;
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
