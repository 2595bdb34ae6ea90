/*
 * tgb_yaml.h -- the YAML front-end of the host API: TensegrityModel builds a tgStructure / tgBuildSpec from the reference's structure
 * files (resources/YamlStructures/...) with the key names, defaults, tag rules and error messages of
 * src/yamlbuilder/TensegrityModel.{h,cpp}.  The reference parses with yaml-cpp; this repo carries its own reader for the YAML subset
 * those files use (block mappings and sequences by indentation, flow sequences, plain / quoted scalars, comments).
 * Not built here: contact-cable builders (std::invalid_argument says so).
 */
#ifndef TGB_YAML_H
#define TGB_YAML_H

#include "tgb_creator.h"

/* A parsed YAML node: null, scalar, sequence or (ordered) mapping. */
class Yam {
public:
    enum Type { Null, Scalar, Sequence, Map };
    Yam() : m_type(Null) {}
    Type type() const { return m_type; }
    bool defined() const { return m_type != Null; }
    operator bool() const { return defined(); }
    std::size_t size() const { return m_type == Sequence ? m_seq.size() : m_type == Map ? m_map.size() : 0; }
    const Yam& operator[](std::size_t i) const;           /* sequence element; out of range -> std::invalid_argument */
    const Yam& operator[](const std::string& key) const;  /* mapping value, or the null node */
    const Yam& operator[](const char* key) const { return (*this)[std::string(key)]; }
    const Yam& operator[](int i) const { return (*this)[(std::size_t)i]; }
    const std::vector<std::pair<std::string, Yam> >& items() const { return m_map; }
    const std::vector<Yam>& elements() const { return m_seq; }
    const std::string& str() const { return m_scalar; }
    double asDouble() const; /* not a number -> std::invalid_argument */
    bool asBool() const;
    static Yam load(const std::string& text);             /* malformed input -> std::invalid_argument */
    static Yam loadFile(const std::string& path);         /* unreadable -> std::runtime_error (yaml-cpp: YAML::BadFile) */
private:
    friend class YamParser;
    Type m_type;
    std::string m_scalar;
    std::vector<Yam> m_seq;
    std::vector<std::pair<std::string, Yam> > m_map;
};

/* src/yamlbuilder/TensegrityModel.h:52-427 */
class TensegrityModel : public tgSubject<TensegrityModel>, public tgModel {
public:
    explicit TensegrityModel(const std::string& structurePath);
    TensegrityModel(const std::string& structurePath, bool debugging);
    virtual ~TensegrityModel();
    virtual void setup(tgWorld& world);
    virtual void step(double dt); /* dt <= 0 -> std::invalid_argument */
    virtual void teardown();
    virtual bool stepsOnHost() const { return numObservers() > 0; }
    const std::vector<tgSpringCableActuator*>& getAllActuators() const { return allActuators; }

    /* defaults of TensegrityModel.h:69-187 */
    static constexpr double rodRadius = 0.5, rodDensity = 1.0, rodFriction = 1.0, rodRollFriction = 0.0, rodRestitution = 0.2;
    static constexpr double boxWidth = 1.0, boxHeight = 1.0, boxDensity = 1.0, boxFriction = 1.0, boxRollFriction = 0.0, boxRestitution = 0.2;
    static constexpr double sphereRadius = 0.5, sphereDensity = 1.0, sphereFriction = 1.0, sphereRollFriction = 0.0, sphereRestitution = 0.2;
    static constexpr double stringStiffness = 1000.0, stringDamping = 10.0, stringPretension = 0.0, stringRadius = 1.0,
                            stringMotorFriction = 0.0, stringMotorInertia = 1.0, stringBackDrivable = 0, stringHistory = 0,
                            stringMaxTension = 1000.0, stringTargetVelocity = 100.0, stringMinActualLength = 0.1,
                            stringMinRestLength = 0.1, stringRotation = 0;
private:
    void buildStructure(tgStructure& structure, const std::string& structurePath, tgBuildSpec& spec);
    void addChildren(tgStructure& structure, const std::string& structurePath, tgBuildSpec& spec, const Yam& children);
    void addChild(tgStructure& structure, const std::string& parentPath, const std::string& childName, const Yam& childStructurePath,
                  tgBuildSpec& spec);
    void addChildRotation(tgStructure& childStructure, const Yam& rotation);
    void addChildScale(tgStructure& childStructure, const Yam& scale);
    void addChildOffset(tgStructure& childStructure, int offsetIndex, const Yam& offset);
    void addChildTranslation(tgStructure& childStructure, const Yam& translation);
    void addNodes(tgStructure& structure, const Yam& nodes);
    void addPairGroups(tgStructure& structure, const Yam& pair_groups);
    void addBondGroups(tgStructure& structure, const Yam& bond_groups, tgBuildSpec& spec);
    void addBonds(tgStructure& structure, const std::string& bonds, const std::string& tags, const Yam& pairs, tgBuildSpec& spec);
    void addNodeNodePairs(tgStructure& structure, const std::string& tags, const Yam& pairs, const std::string* childStructure1Name = 0,
                          const std::string* childStructure2Name = 0);
    void addNodeEdgePairs(tgStructure& structure, const std::string& tags, const Yam& pairs, const std::string& childStructure1Name,
                          const std::string& childStructure2Name, tgBuildSpec& spec);
    void removePair(tgStructure& structure, const tgNode* from, const tgNode* to, bool isEdgePair,
                    const std::vector<tgBuildSpec::RigidAgent>& rigidAgents);
    tgNode& getNode(tgStructure& structure, const std::string& nodePath);
    tgStructure& getStructure(tgStructure& parentStructure, const std::string& structurePath);
    void addBuilders(tgBuildSpec& spec, const Yam& builders);
    void addRodBuilder(const std::string& builderClass, const std::string& tagMatch, const Yam& parameters, tgBuildSpec& spec);
    void addBasicActuatorBuilder(const std::string& builderClass, const std::string& tagMatch, const Yam& parameters, tgBuildSpec& spec);
    void addKinematicActuatorBuilder(const std::string& builderClass, const std::string& tagMatch, const Yam& parameters, tgBuildSpec& spec);
    void addBoxBuilder(const std::string& builderClass, const std::string& tagMatch, const Yam& parameters, tgBuildSpec& spec);
    void addSphereBuilder(const std::string& builderClass, const std::string& tagMatch, const Yam& parameters, tgBuildSpec& spec);
    void yamlContainsOnly(const Yam& yam, const std::string& structurePath, const std::vector<std::string>& keys);
    void yamlNoDuplicates(const Yam& yam, const std::string& structurePath);
    std::string topLvlStructurePath;
    bool debugging_on;
    std::vector<tgSpringCableActuator*> allActuators;
};

#endif /* TGB_YAML_H */
