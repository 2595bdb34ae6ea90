/*
 * tgb_yaml.cpp -- YAML reader for the structure files and TensegrityModel (see tgb_yaml.h for the reference lines).
 */
#include "tgb_yaml.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>

/* ---- Yam ----------------------------------------------------------------------------------------------------------------- */
namespace {
const Yam kNullYam;
std::string trim(const std::string& s) {
    std::size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) ++a;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
    return s.substr(a, b - a);
}
std::string unquote(const std::string& s) {
    if (s.size() >= 2 && ((s[0] == '"' && s[s.size() - 1] == '"') || (s[0] == '\'' && s[s.size() - 1] == '\''))) return s.substr(1, s.size() - 2);
    return s;
}
} // namespace

const Yam& Yam::operator[](std::size_t i) const {
    if (m_type != Sequence || i >= m_seq.size()) throw std::invalid_argument("YAML: sequence index out of range");
    return m_seq[i];
}
const Yam& Yam::operator[](const std::string& key) const {
    if (m_type == Map)
        for (std::size_t i = 0; i < m_map.size(); ++i)
            if (m_map[i].first == key) return m_map[i].second;
    return kNullYam;
}
double Yam::asDouble() const {
    if (m_type != Scalar) throw std::invalid_argument("YAML: bad conversion (expected a number)");
    char* end = 0;
    const double v = std::strtod(m_scalar.c_str(), &end);
    if (end == m_scalar.c_str() || *end != 0) throw std::invalid_argument("YAML: bad conversion of '" + m_scalar + "' to a number");
    return v;
}
bool Yam::asBool() const {
    if (m_type != Scalar) throw std::invalid_argument("YAML: bad conversion (expected a boolean)");
    std::string v = m_scalar;
    std::transform(v.begin(), v.end(), v.begin(), ::tolower);
    if (v == "true" || v == "yes" || v == "on" || v == "y" || v == "1") return true;
    if (v == "false" || v == "no" || v == "off" || v == "n" || v == "0") return false;
    throw std::invalid_argument("YAML: bad conversion of '" + m_scalar + "' to a boolean");
}

/* Indentation-driven reader for the subset of YAML the structure files use. */
class YamParser {
public:
    explicit YamParser(const std::string& text) {
        std::istringstream in(text);
        std::string raw;
        while (std::getline(in, raw)) {
            /* strip comments (a '#' at the start of the line or after white space, outside quotes) */
            bool inS = false, inD = false;
            std::size_t cut = raw.size();
            for (std::size_t i = 0; i < raw.size(); ++i) {
                const char c = raw[i];
                if (c == '\'' && !inD) inS = !inS;
                else if (c == '"' && !inS) inD = !inD;
                else if (c == '#' && !inS && !inD && (i == 0 || raw[i - 1] == ' ' || raw[i - 1] == '\t')) { cut = i; break; }
            }
            std::string line = raw.substr(0, cut);
            if (trim(line).empty() || trim(line) == "---") continue;
            if (line.find('\t') != std::string::npos && line.find_first_not_of(" \t") != std::string::npos &&
                line.substr(0, line.find_first_not_of(" \t")).find('\t') != std::string::npos)
                throw std::invalid_argument("YAML: tabs are not allowed for indentation");
            Line l;
            l.indent = (int)line.find_first_not_of(' ');
            l.text = trim(line);
            m_lines.push_back(l);
        }
        m_pos = 0;
    }
    Yam parse() {
        if (m_lines.empty()) return Yam();
        Yam root = parseBlock(m_lines[0].indent);
        if (m_pos != m_lines.size()) throw std::invalid_argument("YAML: inconsistent indentation near '" + m_lines[m_pos].text + "'");
        return root;
    }
private:
    struct Line { int indent; std::string text; };
    std::vector<Line> m_lines;
    std::size_t m_pos;

    static Yam scalar(const std::string& s) {
        Yam y;
        const std::string v = trim(s);
        if (v.empty() || v == "~" || v == "null") return y;
        y.m_type = Yam::Scalar;
        y.m_scalar = unquote(v);
        return y;
    }
    /* "[a, [b, c], d]" */
    static Yam flow(const std::string& s, std::size_t& i) {
        Yam y;
        y.m_type = Yam::Sequence;
        ++i; /* '[' */
        std::string cur;
        bool haveItem = false;
        for (; i < s.size(); ++i) {
            const char c = s[i];
            if (c == '[') {
                y.m_seq.push_back(flow(s, i));
                haveItem = true;
                cur.clear();
                /* flow() leaves i on the matching ']' */
            } else if (c == ',' || c == ']') {
                if (!trim(cur).empty()) y.m_seq.push_back(scalar(cur));
                else if (!haveItem && c == ',') throw std::invalid_argument("YAML: empty entry in a flow sequence");
                cur.clear();
                haveItem = false;
                if (c == ']') return y;
            } else {
                cur += c;
            }
        }
        throw std::invalid_argument("YAML: unterminated flow sequence");
    }
    static Yam value(const std::string& s) {
        const std::string v = trim(s);
        if (!v.empty() && v[0] == '[') {
            std::size_t i = 0;
            Yam y = flow(v, i);
            if (trim(v.substr(i + 1)).size()) throw std::invalid_argument("YAML: text after a flow sequence");
            return y;
        }
        if (!v.empty() && v[0] == '{') throw std::invalid_argument("YAML: flow mappings are not supported by this reader");
        return scalar(v);
    }
    /* position of the ':' that ends a mapping key, or npos */
    static std::size_t keyEnd(const std::string& t) {
        bool inS = false, inD = false;
        int depth = 0;
        for (std::size_t i = 0; i < t.size(); ++i) {
            const char c = t[i];
            if (c == '\'' && !inD) inS = !inS;
            else if (c == '"' && !inS) inD = !inD;
            else if (!inS && !inD && c == '[') ++depth;
            else if (!inS && !inD && c == ']') --depth;
            else if (!inS && !inD && depth == 0 && c == ':' && (i + 1 == t.size() || t[i + 1] == ' ')) return i;
        }
        return std::string::npos;
    }
    Yam parseBlock(int indent) {
        const Line& first = m_lines[m_pos];
        if (first.text[0] == '-' && (first.text.size() == 1 || first.text[1] == ' ')) return parseSequence(indent);
        if (keyEnd(first.text) != std::string::npos) return parseMap(indent);
        Yam y = value(first.text);
        ++m_pos;
        return y;
    }
    Yam parseSequence(int indent) {
        Yam y;
        y.m_type = Yam::Sequence;
        while (m_pos < m_lines.size() && m_lines[m_pos].indent == indent && m_lines[m_pos].text[0] == '-' &&
               (m_lines[m_pos].text.size() == 1 || m_lines[m_pos].text[1] == ' ')) {
            const std::string rest = trim(m_lines[m_pos].text.substr(1));
            if (rest.empty()) {
                ++m_pos;
                if (m_pos < m_lines.size() && m_lines[m_pos].indent > indent) y.m_seq.push_back(parseBlock(m_lines[m_pos].indent));
                else y.m_seq.push_back(Yam());
            } else if (keyEnd(rest) != std::string::npos && rest[0] != '[') {
                throw std::invalid_argument("YAML: mappings inside block sequences are not supported by this reader");
            } else {
                y.m_seq.push_back(value(rest));
                ++m_pos;
            }
        }
        return y;
    }
    Yam parseMap(int indent) {
        Yam y;
        y.m_type = Yam::Map;
        while (m_pos < m_lines.size() && m_lines[m_pos].indent == indent) {
            const std::string& t = m_lines[m_pos].text;
            const std::size_t k = keyEnd(t);
            if (k == std::string::npos) throw std::invalid_argument("YAML: expected 'key: value' near '" + t + "'");
            const std::string key = unquote(trim(t.substr(0, k)));
            const std::string rest = trim(t.substr(k + 1));
            ++m_pos;
            Yam v;
            if (!rest.empty()) {
                v = value(rest);
            } else if (m_pos < m_lines.size() && (m_lines[m_pos].indent > indent ||
                                                   (m_lines[m_pos].indent == indent && m_lines[m_pos].text[0] == '-' &&
                                                    (m_lines[m_pos].text.size() == 1 || m_lines[m_pos].text[1] == ' ')))) {
                v = parseBlock(m_lines[m_pos].indent); /* a block sequence may sit at its key's indentation */
            }
            y.m_map.push_back(std::make_pair(key, v));
        }
        if (m_pos < m_lines.size() && m_lines[m_pos].indent > indent)
            throw std::invalid_argument("YAML: inconsistent indentation near '" + m_lines[m_pos].text + "'");
        return y;
    }
};

Yam Yam::load(const std::string& text) { return YamParser(text).parse(); }
Yam Yam::loadFile(const std::string& path) {
    std::ifstream in(path.c_str());
    if (!in.is_open()) throw std::runtime_error("bad file: " + path);
    std::stringstream ss;
    ss << in.rdbuf();
    return load(ss.str());
}

/* ---- TensegrityModel ------------------------------------------------------------------------------------------------------ */
TensegrityModel::TensegrityModel(const std::string& structurePath) : topLvlStructurePath(structurePath), debugging_on(false) {}
TensegrityModel::TensegrityModel(const std::string& structurePath, bool debugging) : topLvlStructurePath(structurePath), debugging_on(debugging) {}
TensegrityModel::~TensegrityModel() {}

void TensegrityModel::setup(tgWorld& world) {
    tgBuildSpec spec;
    /* default builders for the tags rod / string / box / sphere, overwritten by the file's own builders (TensegrityModel.cpp:89-95) */
    const Yam emptyYam;
    addRodBuilder("tgRodInfo", "rod", emptyYam, spec);
    addBasicActuatorBuilder("tgBasicActuatorInfo", "string", emptyYam, spec);
    addBoxBuilder("tgBoxInfo", "box", emptyYam, spec);
    addSphereBuilder("tgSphereInfo", "sphere", emptyYam, spec);

    tgStructure structure;
    buildStructure(structure, topLvlStructurePath, spec);
    tgStructureInfo structureInfo(structure, spec);
    structureInfo.buildInto(*this, world);
    allActuators = tgCast::filter<tgModel, tgSpringCableActuator>(getDescendants());
    notifySetup();
    tgModel::setup(world);
}
void TensegrityModel::step(double timeStep) {
    if (timeStep <= 0.0) throw std::invalid_argument("time step is not positive");
    notifyStep(timeStep);
    tgModel::step(timeStep);
}
void TensegrityModel::teardown() {
    notifyTeardown();
    tgModel::teardown();
    allActuators.clear();
}

void TensegrityModel::buildStructure(tgStructure& structure, const std::string& structurePath, tgBuildSpec& spec) {
    Yam root;
    try {
        root = Yam::loadFile(structurePath);
    } catch (const std::runtime_error&) {
        std::cout << std::endl << "The YAML parser threw a BadFile exception when trying to load one of your YAML files. " << std::endl
                  << "The path of the structure that the parser attempted to load is: '" << structurePath << "'. " << std::endl
                  << "Check to be sure that the file exists, and that you didn't spell the path name incorrectly." << std::endl << std::endl;
        throw;
    }
    const char* rootKeys[] = {"nodes", "pair_groups", "builders", "substructures", "bond_groups"};
    yamlContainsOnly(root, structurePath, std::vector<std::string>(rootKeys, rootKeys + 5));
    yamlNoDuplicates(root, structurePath);
    addChildren(structure, structurePath, spec, root["substructures"]);
    addBuilders(spec, root["builders"]);
    addNodes(structure, root["nodes"]);
    addPairGroups(structure, root["pair_groups"]);
    addBondGroups(structure, root["bond_groups"], spec);
}

namespace {
/* "child1/child2/child3" -> names (TensegrityModel.cpp:128-135) */
std::vector<std::string> splitSlash(const std::string& combos) {
    std::vector<std::string> out;
    std::string rest = combos + "/";
    while (rest.find("/") != std::string::npos) {
        out.push_back(rest.substr(0, rest.find("/")));
        rest = rest.substr(rest.find("/") + 1);
    }
    return out;
}
btVector3 vec3(const Yam& y) { return btVector3(y[0].asDouble(), y[1].asDouble(), y[2].asDouble()); }
} // namespace

void TensegrityModel::addChildren(tgStructure& structure, const std::string& structurePath, tgBuildSpec& spec, const Yam& children) {
    if (!children) return;
    const char* keys[] = {"path", "rotation", "translation", "scale", "offset"};
    const std::vector<std::string> allowed(keys, keys + 5);
    const std::vector<std::pair<std::string, Yam> >& items = children.items();
    /* all the children first, then their rotations, then scale / offset / translation (TensegrityModel.cpp:123-167) */
    for (std::size_t c = 0; c < items.size(); ++c) {
        yamlContainsOnly(items[c].second, structurePath, allowed);
        const std::vector<std::string> names = splitSlash(items[c].first);
        for (std::size_t n = 0; n < names.size(); ++n) addChild(structure, structurePath, names[n], items[c].second["path"], spec);
    }
    for (std::size_t c = 0; c < items.size(); ++c) {
        const std::vector<std::string> names = splitSlash(items[c].first);
        for (std::size_t n = 0; n < names.size(); ++n) addChildRotation(structure.findChild(names[n]), items[c].second["rotation"]);
    }
    for (std::size_t c = 0; c < items.size(); ++c) {
        const std::vector<std::string> names = splitSlash(items[c].first);
        for (std::size_t n = 0; n < names.size(); ++n) {
            tgStructure& child = structure.findChild(names[n]);
            addChildScale(child, items[c].second["scale"]);
            addChildOffset(child, (int)n, items[c].second["offset"]);
            addChildTranslation(child, items[c].second["translation"]);
        }
    }
}
void TensegrityModel::addChild(tgStructure& structure, const std::string& parentPath, const std::string& childName, const Yam& childStructurePath,
                               tgBuildSpec& spec) {
    if (!childStructurePath) return;
    std::string childPath = childStructurePath.str();
    if (childPath[0] != '/') childPath = parentPath.substr(0, parentPath.rfind("/") + 1) + childPath; /* relative to the parent file */
    tgStructure childStructure = tgStructure(childName);
    buildStructure(childStructure, childPath, spec);
    structure.addChild(childStructure);
}
void TensegrityModel::addChildRotation(tgStructure& childStructure, const Yam& rotation) {
    if (!rotation) return;
    const Yam& reference = rotation["reference"];
    const Yam& axis = rotation["axis"];
    const Yam& angle = rotation["angle"];
    if (axis && angle) {
        const btVector3 axisVector = vec3(axis);
        const double angleRadians = angle.asDouble() * (M_PI / 180.0); /* tgUtil::deg2rad */
        const btVector3 referenceVector = reference ? vec3(reference) : childStructure.getCentroid();
        childStructure.addRotation(referenceVector, axisVector, angleRadians);
    }
}
void TensegrityModel::addChildScale(tgStructure& childStructure, const Yam& scale) {
    if (!scale) return;
    childStructure.scale(scale.asDouble());
}
void TensegrityModel::addChildOffset(tgStructure& childStructure, int offsetIndex, const Yam& offset) {
    if (!offset) return;
    childStructure.move(vec3(offset) * (btScalar)offsetIndex);
}
void TensegrityModel::addChildTranslation(tgStructure& childStructure, const Yam& translation) {
    if (!translation) return;
    childStructure.move(vec3(translation));
}
void TensegrityModel::addNodes(tgStructure& structure, const Yam& nodes) {
    if (!nodes) return;
    const std::vector<std::pair<std::string, Yam> >& items = nodes.items();
    for (std::size_t i = 0; i < items.size(); ++i) {
        const Yam& xyz = items[i].second;
        structure.addNode(xyz[0].asDouble(), xyz[1].asDouble(), xyz[2].asDouble(), items[i].first);
    }
}
void TensegrityModel::addPairGroups(tgStructure& structure, const Yam& pair_groups) {
    if (!pair_groups) return;
    const std::vector<std::pair<std::string, Yam> >& items = pair_groups.items();
    for (std::size_t i = 0; i < items.size(); ++i) addNodeNodePairs(structure, items[i].first, items[i].second);
}
void TensegrityModel::addBondGroups(tgStructure& structure, const Yam& bond_groups, tgBuildSpec& spec) {
    if (!bond_groups) return;
    const std::vector<std::pair<std::string, Yam> >& first = bond_groups.items();
    for (std::size_t a = 0; a < first.size(); ++a) {
        std::string bonds, tags;
        bool bondsSet = false, tagsSet = false;
        const std::string& firstLvlKey = first[a].first;
        if (firstLvlKey.find("/") == std::string::npos) { tags = firstLvlKey; tagsSet = true; } /* a tag */
        else { bonds = firstLvlKey; bondsSet = true; }                                            /* a bond */
        const std::vector<std::pair<std::string, Yam> >& second = first[a].second.items();
        for (std::size_t b = 0; b < second.size(); ++b) {
            const std::string& secondLvlKey = second[b].first;
            if (secondLvlKey.find("/") == std::string::npos) {
                if (tagsSet) throw std::invalid_argument("Invalid nesting of tags: " + secondLvlKey);
                tags = secondLvlKey;
            } else {
                if (bondsSet) throw std::invalid_argument("Invalid nesting of bonds: " + secondLvlKey);
                bonds = secondLvlKey;
            }
            addBonds(structure, bonds, tags, second[b].second, spec);
        }
    }
}
void TensegrityModel::addBonds(tgStructure& structure, const std::string& bonds, const std::string& tags, const Yam& pairs, tgBuildSpec& spec) {
    if (!pairs) return;
    /* bonds looks like childName1/childName2/.../bondType: consecutive children are bonded pairwise */
    const std::string bondType = bonds.substr(bonds.rfind("/") + 1);
    std::string structureCombos = bonds.substr(0, bonds.rfind("/"));
    if (structureCombos.find("/") == std::string::npos)
        throw std::invalid_argument("Error: all bond groups must specify a bond between 2 or more structures and a bond type");
    while (structureCombos.find("/") != std::string::npos) {
        const std::string childStructure1Name = structureCombos.substr(0, structureCombos.find("/"));
        structureCombos = structureCombos.substr(structureCombos.find("/") + 1);
        const std::string childStructure2Name = structureCombos.substr(0, structureCombos.find("/"));
        if (bondType == "node_node") addNodeNodePairs(structure, tags, pairs, &childStructure1Name, &childStructure2Name);
        else if (bondType == "node_edge") addNodeEdgePairs(structure, tags, pairs, childStructure1Name, childStructure2Name, spec);
        else throw std::invalid_argument("Unsupported bond type: " + bondType);
    }
}
void TensegrityModel::addNodeNodePairs(tgStructure& structure, const std::string& tags, const Yam& pairs, const std::string* childStructure1Name,
                                       const std::string* childStructure2Name) {
    const std::vector<Yam>& list = pairs.elements();
    for (std::size_t i = 0; i < list.size(); ++i) {
        const std::string node1Path = list[i][0].str(), node2Path = list[i][1].str();
        tgNode* node1;
        tgNode* node2;
        std::string pairNewTags = tags;
        if (childStructure1Name && childStructure2Name) {
            node1 = &getNode(structure.findChild(*childStructure1Name), node1Path);
            node2 = &getNode(structure.findChild(*childStructure2Name), node2Path);
            /* three more tags: the two structures the pair connects, and "name1/name2" as in the file (TensegrityModel.cpp:408-415) */
            pairNewTags = tags + " " + *childStructure1Name + " " + *childStructure2Name + " " + *childStructure1Name + "/" + *childStructure2Name;
        } else {
            node1 = &getNode(structure, node1Path);
            node2 = &getNode(structure, node2Path);
        }
        structure.addPair(*node1, *node2, pairNewTags);
    }
}
/* node_edge bonds (TensegrityModel.cpp:429-548): every listed pair names a node of one child (a "ligand") and an edge "a/b" of the
   other (a "receptor").  The second child is turned and moved so that the reference points (nodes, edge midpoints) of both sides
   coincide in centroid, plane normal and first direction; then every receptor edge a-b is replaced by a-node-b, and the strings that
   joined the ligand nodes to each other are dropped. */
namespace {
struct Attachment {
    std::vector<btVector3> refPoints;
};
btVector3 centroidOf(const std::vector<btVector3>& pts) {
    btVector3 c(0, 0, 0);
    for (std::size_t i = 0; i < pts.size(); ++i) c += pts[i];
    return c / (btScalar)pts.size(); /* tgUtil::getCentroid, tgUtil.h:246-254 */
}
btVector3 planeNormal(const std::vector<btVector3>& p) { return ((p[1] - p[0]).cross(p[2] - p[0])).normalize(); }
} // namespace

void TensegrityModel::addNodeEdgePairs(tgStructure& structure, const std::string& tags, const Yam& pairs, const std::string& childStructure1Name,
                                       const std::string& childStructure2Name, tgBuildSpec& spec) {
    if (pairs.size() < 3) throw std::invalid_argument("Error: node_edge bonds must specify at least 3 node_edge pairs");
    tgStructure& child1 = structure.findChild(childStructure1Name);
    tgStructure& child2 = structure.findChild(childStructure2Name);

    Attachment side[2];
    std::vector<tgNode*> ligands;
    std::vector<std::pair<tgNode*, tgNode*> > receptors;
    const std::vector<Yam>& list = pairs.elements();
    for (std::size_t i = 0; i < list.size(); ++i) {
        const std::string where[2] = {list[i][0].str(), list[i][1].str()};
        const bool edge0 = where[0].find("/") != std::string::npos, edge1 = where[1].find("/") != std::string::npos;
        if (edge0 == edge1) throw std::invalid_argument("Invalid node_edge bond: " + where[0] + ", " + where[1]);
        for (int s = 0; s < 2; ++s) {
            tgStructure& child = s == 0 ? child1 : child2;
            const std::size_t slash = where[s].find("/");
            if (slash == std::string::npos) {
                tgNode* n = &getNode(child, where[s]);
                ligands.push_back(n);
                side[s].refPoints.push_back(*n);
            } else {
                tgNode* a = &getNode(child, where[s].substr(0, slash));
                tgNode* b = &getNode(child, where[s].substr(slash + 1));
                receptors.push_back(std::make_pair(a, b));
                side[s].refPoints.push_back((*a + *b) / 2);
            }
        }
    }

    /* align child 2 with child 1 (TensegrityModel.cpp:497-528; only the first reference point and the centroid of side 2 are
       tracked through the motion, like there) */
    const btVector3 c1 = centroidOf(side[0].refPoints);
    btVector3 c2 = centroidOf(side[1].refPoints);
    const btVector3 n1 = planeNormal(side[0].refPoints), n2 = planeNormal(side[1].refPoints);
    const btVector3 fallBackAxis = (side[1].refPoints[1] - side[1].refPoints[0]).normalize();
    const btQuaternion alignNormals = tgUtil::getQuaternionBetween(n2, n1, fallBackAxis);
    btVector3 p2 = side[1].refPoints[0];
    child2.addRotation(c2, alignNormals);
    tgUtil::addRotation(p2, c2, alignNormals);
    tgUtil::addRotation(c2, c2, alignNormals);
    const btVector3 shift = c1 - c2;
    child2.move(shift);
    p2 += shift;
    c2 += shift;
    child2.addRotation(c1, tgUtil::getQuaternionBetween(p2 - c1, side[0].refPoints[0] - c1, n1));

    const std::vector<tgBuildSpec::RigidAgent>& rigidAgents = spec.getRigidAgents();
    for (std::size_t i = 0; i < ligands.size(); ++i) {
        /* which child holds a pair is not recorded: try both */
        removePair(child1, receptors[i].first, receptors[i].second, true, rigidAgents);
        removePair(child2, receptors[i].first, receptors[i].second, true, rigidAgents);
        for (std::size_t j = 0; j < ligands.size(); ++j) {
            removePair(child1, ligands[i], ligands[j], false, rigidAgents);
            removePair(child2, ligands[i], ligands[j], false, rigidAgents);
        }
        structure.addPair(*receptors[i].first, *ligands[i], tags);
        structure.addPair(*ligands[i], *receptors[i].second, tags);
    }
}
void TensegrityModel::removePair(tgStructure& structure, const tgNode* from, const tgNode* to, bool isEdgePair,
                                 const std::vector<tgBuildSpec::RigidAgent>& rigidAgents) {
    tgPair* pair;
    try {
        pair = &structure.findPair(*from, *to);
    } catch (const std::invalid_argument&) {
        return;
    }
    for (std::size_t i = 0; i < rigidAgents.size(); ++i)
        if (rigidAgents[i].tagSearch.matches(pair->getTags())) {
            /* rods stay; an edge that is a rod cannot take a node */
            if (isEdgePair) throw std::invalid_argument("Edges in node_edge bonds cannot be rods");
            return;
        }
    structure.removePair(*pair);
}

tgNode& TensegrityModel::getNode(tgStructure& structure, const std::string& nodePath) {
    /* nodePath looks like 'parentStructure.childStructure.nodeName' */
    if (nodePath.find(".") == std::string::npos) return structure.findNode(nodePath);
    const std::string structurePath = nodePath.substr(0, nodePath.rfind("."));
    const std::string nodeName = nodePath.substr(nodePath.rfind(".") + 1);
    return getStructure(structure, structurePath).findNode(nodeName);
}
tgStructure& TensegrityModel::getStructure(tgStructure& parentStructure, const std::string& childStructurePath) {
    if (childStructurePath.find(".") == std::string::npos) return parentStructure.findChild(childStructurePath);
    const std::string childStructureName = childStructurePath.substr(0, childStructurePath.rfind("."));
    const std::string remaining = childStructurePath.substr(childStructurePath.rfind(".") + 1);
    return getStructure(parentStructure.findChild(childStructureName), remaining);
}

void TensegrityModel::addBuilders(tgBuildSpec& spec, const Yam& builders) {
    const std::vector<std::pair<std::string, Yam> >& items = builders.items();
    for (std::size_t i = 0; i < items.size(); ++i) {
        const std::string& tagMatch = items[i].first;
        if (!items[i].second["class"]) throw std::invalid_argument("Builder class not supplied for tag: " + tagMatch);
        const std::string builderClass = items[i].second["class"].str();
        const Yam& parameters = items[i].second["parameters"];
        if (builderClass == "tgRodInfo") addRodBuilder(builderClass, tagMatch, parameters, spec);
        else if (builderClass == "tgBasicActuatorInfo" || builderClass == "tgBasicContactCableInfo")
            addBasicActuatorBuilder(builderClass, tagMatch, parameters, spec);
        else if (builderClass == "tgKinematicActuatorInfo") addKinematicActuatorBuilder(builderClass, tagMatch, parameters, spec);
        else if (builderClass == "tgSphereInfo") addSphereBuilder(builderClass, tagMatch, parameters, spec);
        else if (builderClass == "tgBoxInfo") addBoxBuilder(builderClass, tagMatch, parameters, spec);
        else if (builderClass == "tgKinematicContactCableInfo")
            throw std::invalid_argument("Unsupported builder class in this build: " + builderClass);
        else throw std::invalid_argument("Unsupported builder class: " + builderClass);
    }
}
namespace {
/* defaults overwritten by the file's parameters; an unknown name is an error (TensegrityModel.cpp:637-646) */
void overwrite(std::map<std::string, double>& p, const Yam& parameters, const std::string& builderClass, std::map<std::string, bool>* booleans = 0) {
    if (!parameters) return;
    const std::vector<std::pair<std::string, Yam> >& items = parameters.items();
    for (std::size_t i = 0; i < items.size(); ++i) {
        const std::string& name = items[i].first;
        if (p.find(name) != p.end()) p[name] = items[i].second.asDouble();
        else if (booleans && booleans->find(name) != booleans->end()) (*booleans)[name] = items[i].second.asBool();
        else throw std::invalid_argument("Unsupported " + builderClass + " parameter: " + name);
    }
}
} // namespace
void TensegrityModel::addRodBuilder(const std::string& builderClass, const std::string& tagMatch, const Yam& parameters, tgBuildSpec& spec) {
    std::map<std::string, double> rp;
    rp["radius"] = rodRadius; rp["density"] = rodDensity; rp["friction"] = rodFriction; rp["roll_friction"] = rodRollFriction;
    rp["restitution"] = rodRestitution;
    overwrite(rp, parameters, builderClass);
    spec.addBuilder(tagMatch, new tgRodInfo(tgRod::Config(rp["radius"], rp["density"], rp["friction"], rp["roll_friction"], rp["restitution"])));
}
void TensegrityModel::addBasicActuatorBuilder(const std::string& builderClass, const std::string& tagMatch, const Yam& parameters, tgBuildSpec& spec) {
    std::map<std::string, double> d;
    std::map<std::string, bool> b;
    d["stiffness"] = stringStiffness; d["damping"] = stringDamping; d["pretension"] = stringPretension; d["history"] = stringHistory;
    d["max_tension"] = stringMaxTension; d["target_velocity"] = stringTargetVelocity; d["min_actual_length"] = stringMinActualLength;
    d["min_rest_length"] = stringMinRestLength; d["rotation"] = stringRotation;
    b["moveCablePointAToEdge"] = true; b["moveCablePointBToEdge"] = true;
    overwrite(d, parameters, builderClass, &b);
    const tgBasicActuator::Config config(d["stiffness"], d["damping"], d["pretension"], d["history"] != 0.0, d["max_tension"], d["target_velocity"],
                                         d["min_actual_length"], d["min_rest_length"], d["rotation"], b["moveCablePointAToEdge"],
                                         b["moveCablePointBToEdge"]);
    /* both builders share tgBasicActuator::Config (TensegrityModel.cpp:726-734) */
    if (builderClass == "tgBasicContactCableInfo") spec.addBuilder(tagMatch, new tgBasicContactCableInfo(config));
    else spec.addBuilder(tagMatch, new tgBasicActuatorInfo(config));
}
void TensegrityModel::addKinematicActuatorBuilder(const std::string& builderClass, const std::string& tagMatch, const Yam& parameters, tgBuildSpec& spec) {
    std::map<std::string, double> k;
    k["stiffness"] = stringStiffness; k["damping"] = stringDamping; k["pretension"] = stringPretension; k["radius"] = stringRadius;
    k["motor_friction"] = stringMotorFriction; k["motor_inertia"] = stringMotorInertia; k["back_drivable"] = stringBackDrivable;
    k["history"] = stringHistory; k["max_tension"] = stringMaxTension; k["target_velocity"] = stringTargetVelocity;
    k["min_actual_length"] = stringMinActualLength; k["min_rest_length"] = stringMinRestLength; k["rotation"] = stringRotation;
    overwrite(k, parameters, builderClass);
    spec.addBuilder(tagMatch, new tgKinematicActuatorInfo(tgKinematicActuator::Config(k["stiffness"], k["damping"], k["pretension"], k["radius"],
                                                                                       k["motor_friction"], k["motor_inertia"], k["back_drivable"] != 0.0,
                                                                                       k["history"] != 0.0, k["max_tension"], k["target_velocity"],
                                                                                       k["min_actual_length"], k["min_rest_length"], k["rotation"])));
}
void TensegrityModel::addBoxBuilder(const std::string& builderClass, const std::string& tagMatch, const Yam& parameters, tgBuildSpec& spec) {
    std::map<std::string, double> bp;
    bp["width"] = boxWidth; bp["height"] = boxHeight; bp["density"] = boxDensity; bp["friction"] = boxFriction;
    bp["roll_friction"] = boxRollFriction; bp["restitution"] = boxRestitution;
    overwrite(bp, parameters, builderClass);
    spec.addBuilder(tagMatch, new tgBoxInfo(tgBox::Config(bp["width"], bp["height"], bp["density"], bp["friction"], bp["roll_friction"], bp["restitution"])));
}
void TensegrityModel::addSphereBuilder(const std::string& builderClass, const std::string& tagMatch, const Yam& parameters, tgBuildSpec& spec) {
    std::map<std::string, double> sp;
    sp["radius"] = sphereRadius; sp["density"] = sphereDensity; sp["friction"] = sphereFriction; sp["roll_friction"] = sphereRollFriction;
    sp["restitution"] = sphereRestitution;
    overwrite(sp, parameters, builderClass);
    spec.addBuilder(tagMatch, new tgSphereInfo(tgSphere::Config(sp["radius"], sp["density"], sp["friction"], sp["roll_friction"], sp["restitution"])));
}
void TensegrityModel::yamlNoDuplicates(const Yam& yam, const std::string& structurePath) {
    std::set<std::string> keys;
    const std::vector<std::pair<std::string, Yam> >& items = yam.items();
    for (std::size_t i = 0; i < items.size(); ++i) {
        if (items[i].second.type() == Yam::Map) yamlNoDuplicates(items[i].second, structurePath);
        if (!keys.insert(items[i].first).second)
            throw std::invalid_argument(structurePath.substr(structurePath.rfind("/") + 1) + " contains duplicate key: " + items[i].first);
    }
}
void TensegrityModel::yamlContainsOnly(const Yam& yam, const std::string& structurePath, const std::vector<std::string>& keys) {
    const std::vector<std::pair<std::string, Yam> >& items = yam.items();
    for (std::size_t i = 0; i < items.size(); ++i)
        if (std::find(keys.begin(), keys.end(), items[i].first) == keys.end())
            throw std::invalid_argument(structurePath.substr(structurePath.rfind("/") + 1) + " contains invalid key: " + items[i].first);
}
