"""YAML front-end of the host API (first half): `TensegrityModel` builds a structure file of the reference's
format (src/yamlbuilder/TensegrityModel.cpp) into the same flat descriptors the C++ example models produce.  The structure files are
written by these tests (the GPU box has no /root/reference); when the reference tree is present the shipped
resources/YamlStructures files are loaded as well."""
import ctypes as C
import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ntrtsim_b200 as nt  # noqa: E402
from ntrtsim_b200 import examples  # noqa: E402

REF_YAML = "/root/reference/resources/YamlStructures"

PRISM = """
# the 3-prism of src/examples/3_prism/PrismModel.cpp (edge 10, width 10, height 20, lifted by 10)
nodes:
  b0: [-5, 10, 0]
  b1: [5, 10, 0]
  b2: [0, 10, 10]
  t0: [-5, 30, 0]
  t1: [5, 30, 0]
  t2: [0, 30, 10]

pair_groups:
  rod:
    - [b0, t1]
    - [b1, t2]
    - [b2, t0]
  muscle:
    - [b0, b1]
    - [b1, b2]
    - [b2, b0]
    - [t0, t1]
    - [t1, t2]
    - [t2, t0]
    - [b0, t0]
    - [b1, t1]
    - [b2, t2]

builders:
  rod:
    class: tgRodInfo
    parameters:
      density: 0.2
      radius: 0.31
  muscle:
    class: tgBasicActuatorInfo
    parameters:
      stiffness: 1000
      damping: 10
      pretension: 500
"""

# a regular 3-prism around the y axis whose default-tagged rods / strings pick up the builders of the parent file
BASE_PRISM = """
nodes:
  bottom1: [4, 0, 0]
  bottom2: [-2, 0, 3.4641016151377544]
  bottom3: [-2, 0, -3.4641016151377544]
  top1: [-2, 8, 3.4641016151377544]
  top2: [-2, 8, -3.4641016151377544]
  top3: [4, 8, 0]
pair_groups:
  prism_rod:
    - [bottom1, top1]
    - [bottom2, top2]
    - [bottom3, top3]
  prism_string:
    - [bottom1, bottom2]
    - [bottom2, bottom3]
    - [bottom3, bottom1]
    - [top1, top2]
    - [top2, top3]
    - [top3, top1]
    - [bottom1, top3]
    - [bottom2, top1]
    - [bottom3, top2]
builders:
  prism_rod:
    class: tgRodInfo
    parameters:
      density: 0.5
      radius: 0.2
  prism_string:
    class: tgBasicActuatorInfo
    parameters:
      stiffness: 800
      pretension: 100
"""

TOWER = """
substructures:
  p1/p2/p3:
    path: base/Prism.yaml
    offset: [0, 12, 0]
    translation: [0, 3, 0]
bond_groups:
  prism_string:
    p1/p2/p3/node_edge:
      - [top1, bottom1/bottom2]
      - [top2, bottom2/bottom3]
      - [top3, bottom3/bottom1]
"""


def write(tmp_path, name, text):
    p = tmp_path / name
    p.parent.mkdir(parents=True, exist_ok=True)
    p.write_text(text)
    return str(p)


def desc_array(items, fields):
    return np.array([[float(np.ravel(getattr(it, f))[k]) for f in fields for k in range(np.size(getattr(it, f)))] for it in items])


def flat(model):
    """All numbers of a FlatModel's descriptors as arrays (field order of the ctypes structs)."""
    def rec(items):
        rows = []
        for it in items:
            row = []
            for name, _ in it._fields_:
                v = getattr(it, name)
                row.extend(list(v) if hasattr(v, "__len__") else [v])
            rows.append([float(x) for x in row])
        return np.array(rows)
    return rec(model.bodies), rec(model.shapes), rec(model.cables)


# ------------------------------------------------------------------------------------------------ the reader
@pytest.mark.parametrize("text", [PRISM, BASE_PRISM, TOWER, """
a: 1
b:
- [x, y/z]
- [ "q r" , 's' ]
c:
  d: [1.5, -2e3, [3, 4]]   # trailing comment
  e: ~
'k l': "v # not a comment"
"""])
def test_reader_agrees_with_a_full_yaml_parser(text):
    yaml = pytest.importorskip("yaml")
    want = yaml.load(text, Loader=yaml.BaseLoader)

    def norm(v):  # BaseLoader keeps '~' and '' as strings; the reader calls them null
        if isinstance(v, dict):
            return {k: norm(x) for k, x in v.items()}
        if isinstance(v, list):
            return [norm(x) for x in v]
        return None if v in ("~", "", "null") else v
    assert json.loads(examples.yaml_to_json(text)) == norm(want)


@pytest.mark.parametrize("text,msg", [("a: [1, 2", "unterminated"), ("a:\n\tb: 1", "tabs"), ("a: {b: 1}", "flow mappings"),
                                      ("a:\n    b: 1\n  c: 2", "indentation")])
def test_reader_rejects_malformed_input(text, msg):
    with pytest.raises(ValueError, match=msg):
        examples.yaml_to_json(text)


# ------------------------------------------------------------------------------------------------ TensegrityModel
def test_yaml_prism_equals_the_cpp_prism_model(tmp_path):
    """The same nodes, pairs and builder parameters through the YAML front-end and through PrismModel::setup give the same bodies,
    collision shapes and cables, bit for bit."""
    path = write(tmp_path, "prism.yaml", PRISM)
    y = examples.build_yaml(path, gravity=981.0)
    c = nt.build_example("3_prism")
    for a, b in zip(flat(y), flat(c)):
        assert a.shape == b.shape and np.array_equal(a, b)
    assert y.gravity == c.gravity and y.dt == c.dt
    assert examples.tree_counts("yaml:" + path) == (3, 9)


def test_builder_defaults_and_world_of_the_yamlbuilder_app(tmp_path):
    """No builders section: pairs tagged rod / string / and nodes tagged sphere get the defaults of TensegrityModel.h:69-187."""
    path = write(tmp_path, "defaults.yaml", """
nodes:
  a: [0, 5, 0]
  b: [0, 5, 6]
  c: [4, 9, 0]
  d: [4, 9, 6]
pair_groups:
  rod:
    - [a, b]
    - [c, d]
  string:
    - [a, c]
    - [b, d]
""")
    m = examples.build_yaml(path)
    assert (m.n_bodies, m.n_cables) == (2, 2) and m.gravity == 98.1 and m.dt == 1e-3
    mass = np.float32(np.pi * 0.5 ** 2 * 6.0 * 1.0)  # density 1, radius 0.5; the reference rounds the mass to float
    assert all(abs(b.mass - mass) < 1e-6 * mass for b in m.bodies)
    for c in m.cables:
        assert (c.stiffness, c.damping, c.max_tension, c.target_velocity, c.min_rest_length) == (1000.0, 10.0, 1000.0, 100.0, 0.1)


def test_substructures_offsets_and_node_edge_bonds(tmp_path):
    """Three copies of a prism stacked by `offset`, joined by node_edge bonds (TensegrityModel.cpp:429-548): each bond replaces the
    three receptor edges of one prism and the three strings between the ligand nodes of the other by six node-to-edge strings, after
    turning the next prism so that its bottom edge midpoints meet the top nodes."""
    write(tmp_path, "base/Prism.yaml", BASE_PRISM)
    path = write(tmp_path, "tower.yaml", TOWER)
    m = examples.build_yaml(path)
    assert (m.n_bodies, m.n_cables) == (9, 3 * 9 - 2 * 6 + 2 * 6)
    assert examples.tree_counts("yaml:" + path) == (9, 27)
    rods = m.initial_rod_state()
    # the bonded prisms end up interleaved (midpoints of the next bottom triangle on the top nodes), not 12 apart as `offset` put them
    ys = np.sort(rods[:, 1])
    assert np.allclose(ys[0:3], 3 + 4) and np.allclose(ys[3:6], 3 + 8 + 4, atol=1e-9) and np.allclose(ys[6:9], 3 + 16 + 4, atol=1e-9)
    # every bond string is taut and shorter than the edge it replaced
    cab = m.initial_cable_state()
    assert (cab[:, 0] > 0).all()


def test_tags_of_node_node_bonds_and_child_paths(tmp_path):
    write(tmp_path, "base/Prism.yaml", BASE_PRISM)
    path = write(tmp_path, "pair.yaml", """
substructures:
  left:
    path: base/Prism.yaml
  right:
    path: base/Prism.yaml
    translation: [12, 0, 0]
    rotation:
      axis: [0, 1, 0]
      angle: 180
bond_groups:
  prism_string:
    left/right/node_node:
      - [bottom1, bottom1]
      - [top3, top3]
""")
    m = examples.build_yaml(path)
    assert (m.n_bodies, m.n_cables) == (6, 20)
    # right was turned half a turn about its own centroid (0, 4, 0) and moved by 12: its bottom1 (4,0,0) -> (-4,0,0) -> (8,0,0), 4 away
    # from left's bottom1; the strings end on the rod surfaces (radius 0.2) and are pretensioned by 100 / 800
    # the parent's own connectors come before those of its children (tgStructureInfo.cpp:317-352)
    for c in m.cables[:2]:
        assert c.body_a < 3 <= c.body_b
        assert 4 - 2 * 0.2 - 0.125 - 1e-9 <= c.rest_length <= 4 - 0.125 + 1e-9


BOX_CART = """
# a box body hung between two rods by four strings; boxes and rods sharing no nodes stay separate bodies
nodes:
  a0: [-3, 4, -2]
  a1: [-3, 4, 2]
  c0: [3, 4, -2]
  c1: [3, 4, 2]
  p0: [0, 2.6, -1.5]
  p1: [0, 2.6, 1.5]
pair_groups:
  rod:
    - [a0, a1]
    - [c0, c1]
  box:
    - [p0, p1]
  string:
    - [a0, p0]
    - [a1, p1]
    - [c0, p0]
    - [c1, p1]
builders:
  rod:
    class: tgRodInfo
    parameters:
      radius: 0.3
      density: 0.4
  box:
    class: tgBoxInfo
    parameters:
      width: 0.8
      height: 0.5
      density: 0.3
      friction: 0.7
      restitution: 0.1
  string:
    class: tgBasicActuatorInfo
    parameters:
      stiffness: 500
      pretension: 50
"""


OBSTACLE_SCENE = """
# a static block (density 0: tgBoxInfo like the obstacles of src/models/obstacles) lying on the ground, y 1.5 .. 3.5, x -3 .. 3, z -1 .. 1;
# above it a rod parallel to the block, a ball, and a short rod that will hit the block's side face
nodes:
  k0: [-3, 2.5, 0]
  k1: [3, 2.5, 0]
  r0: [-2, 4.0, 0.3]
  r1: [2, 4.1, 0.3]
  ball: [0.5, 4.6, -0.4]
  s0: [0, 2.4, 1.9]
  s1: [0, 3.4, 3.2]
pair_groups:
  block:
    - [k0, k1]
  rod:
    - [r0, r1]
    - [s0, s1]
builders:
  block:
    class: tgBoxInfo
    parameters:
      width: 1.0
      height: 1.0
      density: 0
      friction: 0.8
      restitution: 0.0
  rod:
    class: tgRodInfo
    parameters:
      radius: 0.2
      density: 0.5
      friction: 0.9
      restitution: 0.0
  ball:
    class: tgSphereInfo
    parameters:
      radius: 0.5
      density: 0.5
      friction: 0.9
      restitution: 0.0
"""


def test_box_builder(tmp_path):
    """tgBoxInfo: btBoxShape(width, length / 2, height) -- width and height are half extents --, mass of the full box, box inertia of
    the half extents with margin (tgBoxInfo.cpp:100-133)."""
    m = examples.build_yaml(write(tmp_path, "cart.yaml", BOX_CART))
    assert (m.n_bodies, m.n_cables) == (3, 4)
    box = [i for i, b in enumerate(m.bodies) if m.shapes[b.shape_begin].kind == 2]
    assert len(box) == 1
    bi = box[0]
    b = m.bodies[bi]
    s = m.shapes[b.shape_begin]
    assert (s.radius, s.half_length, s.half_extent_z, s.margin) == (0.8, 1.5, 0.5, 0.04)
    mass = float(np.float32(3.0 * 1.6 * 1.0 * 0.3))
    assert abs(b.mass - mass) < 1e-12
    lx, ly, lz = 1.6, 3.0, 1.0
    want = np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly]) * mass / 12.0
    assert np.allclose(list(b.inertia), want, rtol=1e-12)
    assert (b.friction, b.restitution) == (0.7, 0.1)
    # the strings end `width` off the box axis, on the side of the rods (tgBoxInfo::getConnectionPoint)
    for c in m.cables:
        end = np.array(c.local_b if c.body_b == bi else c.local_a)
        assert abs(np.hypot(end[0], end[2]) - 0.8) < 1e-9


@pytest.mark.parametrize("text,msg", [
    ("nodez:\n  a: [0, 0, 0]\n", "contains invalid key: nodez"),
    ("nodes:\n  a: [0, 0, 0]\n  a: [1, 0, 0]\n", "contains duplicate key: a"),
    ("builders:\n  rod:\n    class: tgKinematicContactCableInfo\n", "Unsupported builder class in this build"),
    ("builders:\n  rod:\n    class: tgFooInfo\n", "Unsupported builder class: tgFooInfo"),
    ("builders:\n  rod:\n    parameters:\n      radius: 1\n", "Builder class not supplied for tag: rod"),
    ("builders:\n  rod:\n    class: tgRodInfo\n    parameters:\n      radios: 1\n", "Unsupported tgRodInfo parameter: radios"),
    ("nodes:\n  a: [0, 0, 0]\npair_groups:\n  rod:\n    - [a, b]\n", "Node not found: b"),
    ("substructures:\n  x/y:\n    path: missing.yaml\n", "bad file"),
])
def test_errors_carry_the_reference_messages(tmp_path, text, msg):
    path = write(tmp_path, "bad.yaml", text)
    with pytest.raises(ValueError, match=msg):
        examples.build_yaml(path)


def test_node_edge_bond_needs_three_pairs_and_non_rod_edges(tmp_path):
    write(tmp_path, "base/Prism.yaml", BASE_PRISM)
    two = write(tmp_path, "two.yaml", TOWER.replace("      - [top3, bottom3/bottom1]\n", ""))
    with pytest.raises(ValueError, match="at least 3 node_edge pairs"):
        examples.build_yaml(two)
    rod_edge = write(tmp_path, "rod_edge.yaml", TOWER.replace("bottom1/bottom2", "bottom1/top1"))
    with pytest.raises(ValueError, match="Edges in node_edge bonds cannot be rods"):
        examples.build_yaml(rod_edge)
    both_nodes = write(tmp_path, "nodes.yaml", TOWER.replace("bottom1/bottom2", "bottom1"))
    with pytest.raises(ValueError, match="Invalid node_edge bond"):
        examples.build_yaml(both_nodes)


@pytest.mark.skipif(not os.path.isdir(REF_YAML), reason="reference tree not present")
def test_shipped_structure_files_load():
    """Every file under resources/YamlStructures builds, except 5PrismSpine (its node_edge strings are shorter than pretension /
    stiffness = 3, which the reference's tgSpringCable constructor rejects as well: tgSpringCable.cpp:54-57) .  Body / cable counts of
    the ones with a known composition are checked."""
    want = {"BaseStructures/3Prism.yaml": (3, 9), "BaseStructures/SuperBall.yaml": (6, 24), "Double3Prism.yaml": (6, 18),
            "Spines/3PrismSpine.yaml": (18, 54), "Spines/4PrismSpine.yaml": (20, 60), "Spines/SuperBallSpine.yaml": (12, 48),
            "Spines/TetrahedralSpine.yaml": (6, 40), "Spider.yaml": (30, 96)}
    seen = 0
    for f in sorted(glob.glob(REF_YAML + "/**/*.yaml", recursive=True)):
        rel = os.path.relpath(f, REF_YAML)
        if rel == "Spines/5PrismSpine.yaml":
            with pytest.raises(ValueError, match="Pretension causes string to shorten past rest length"):
                examples.build_example("yaml:" + f)
            continue
        m = examples.build_example("yaml:" + f)
        seen += 1
        if rel in want:
            assert (m.n_bodies, m.n_cables) == want[rel], rel
    assert seen >= 27


def test_scale_kinematic_actuators_and_spheres(tmp_path):
    """`scale` of a substructure (about its own centroid), tgKinematicActuatorInfo parameters and tgSphereInfo at tagged nodes."""
    write(tmp_path, "base/Prism.yaml", BASE_PRISM)
    path = write(tmp_path, "scaled.yaml", """
substructures:
  big:
    path: base/Prism.yaml
    scale: 2
    translation: [0, 10, 0]
nodes:
  weight: [0, 30, 0]
pair_groups:
  motor:
    - [weight, big.top1]
builders:
  weight:
    class: tgSphereInfo
    parameters:
      radius: 0.75
      density: 2.0
  motor:
    class: tgKinematicActuatorInfo
    parameters:
      stiffness: 1500
      damping: 20
      radius: 0.05
      motor_inertia: 0.8
      motor_friction: 0.1
      back_drivable: 1
      max_tension: 900
      target_velocity: 50
""")
    m = examples.build_yaml(path)
    assert (m.n_bodies, m.n_cables) == (4, 10)
    rods = m.initial_rod_state()
    lengths = sorted(2 * m.shapes[b.shape_begin].half_length for b in m.bodies if m.shapes[b.shape_begin].kind == 0)
    base = np.sqrt(6.0 ** 2 + 8.0 ** 2 + 3.4641016151377544 ** 2)  # bottom1 (4,0,0) -> top1 (-2,8,3.46)
    assert np.allclose(lengths, 2 * base)
    ball = [b for b in m.bodies if m.shapes[b.shape_begin].kind == 1][0]
    assert abs(ball.mass - float(np.float32(4.0 / 3.0 * np.pi * 0.75 ** 3 * 2.0))) < 1e-9
    motor = [c for c in m.cables if c.actuator == 2]
    assert len(motor) == 1
    c = motor[0]
    assert (c.stiffness, c.damping, c.motor_radius, c.motor_inertia, c.motor_friction, c.backdrivable, c.max_tension, c.target_velocity) == \
        (1500.0, 20.0, 0.05, 0.8, 0.1, 1, 900.0, 50.0)
    # the scaled prism is centred where the unscaled one was (centroid (0, 4, 0)) before the translation by 10
    assert np.allclose(rods[[i for i, b in enumerate(m.bodies) if m.shapes[b.shape_begin].kind == 0], 0:3].mean(axis=0), [0, 14, 0], atol=1e-9)
