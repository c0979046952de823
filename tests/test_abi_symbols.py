"""CPU-side checks: the C-ABI library builds for sm_100a, loads, and exports every symbol the
header declares; host-side validation works without a GPU (no compute calls)."""
import ctypes as C
import os
import re

import swarm_b200
from swarm_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    header = open(os.path.join(ROOT, "include", "swarm_b200.h")).read()
    declared = set(re.findall(r"\b(swarm_[a-z_]+)\s*\(", header))
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.swarm_abi_version() == _abi.ABI_VERSION


def test_struct_layouts_match_header_field_order():
    header = open(os.path.join(ROOT, "include", "swarm_b200.h")).read()
    body = header[header.index("typedef struct SwarmBuffers {"):header.index("} SwarmBuffers;")]
    names = re.findall(r"\*\s*([a-z_0-9]+)\s*;", body)
    assert tuple(names) == _abi.BUFFER_FIELDS
    body = header[header.index("typedef struct SwarmHostOut {"):header.index("} SwarmHostOut;")]
    assert tuple(re.findall(r"\*\s*([a-z_0-9]+)\s*;", body)) == _abi.HOST_OUT_FIELDS + ("block_host", "block_dev", "flags")
    assert [n for n, _ in _abi.SwarmHostOut._fields_][-4:] == ["block_host", "block_dev", "block_bytes", "flags"]
    bits = dict(re.findall(r"SWARM_FLAG_([A-Z_]+) = (\d+)", header))
    assert [int(bits[n.upper()]) for n in _abi.FLAG_FIELDS] == [1, 2, 4, 8, 16]
    body = header[header.index("typedef struct SwarmSizes {"):header.index("} SwarmSizes;")]
    assert tuple(re.findall(r"int64_t\s+([a-z_0-9]+)\s*;", body)) == tuple(n for n, _ in _abi.SwarmSizes._fields_)


def _cfg(**kw):
    c = _abi.SwarmConfig()
    c.abi_version, c.env_kind, c.num_envs, c.num_drones = _abi.ABI_VERSION, _abi.KIND_SWARM, 16, 8
    c.num_obstacles, c.sensed_obstacles, c.neighbor_k, c.max_steps, c.device = 4, 4, 3, 400, -1
    d = swarm_b200.DroneEnvConfig()
    for n in _abi._DOUBLES:
        setattr(c, n, float(getattr(d, n)))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def test_query_sizes_and_validation_without_gpu():
    lib = _abi.load()
    sz = _abi.SwarmSizes()
    assert lib.swarm_query_sizes(C.byref(_cfg()), C.byref(sz)) == 0
    assert sz.obs_dim == 9 + 12 + 16 and sz.state_dim == 51 and sz.obs == 16 * 8 * 37
    assert sz.pos4 == 16 * 8 * 4 and sz.rng == 64 and sz.actions == 16 * 8 * 3
    c = _cfg(env_kind=_abi.KIND_SINGLE, num_drones=1)
    assert lib.swarm_query_sizes(C.byref(c), C.byref(sz)) == 0 and sz.obs_dim == 25
    for bad in (dict(abi_version=99), dict(num_envs=0), dict(num_drones=0), dict(num_drones=129),
                dict(neighbor_k=9), dict(sensed_obstacles=-1), dict(env_kind=7), dict(norm_mode=2),
                dict(env_kind=_abi.KIND_SINGLE, num_drones=2)):
        rc = lib.swarm_query_sizes(C.byref(_cfg(**bad)), C.byref(sz))
        assert rc < 0, bad
        assert lib.swarm_last_error()


def test_config_mirror_matches_reference_defaults_and_drops_unknown_keys():
    c = swarm_b200.DroneEnvConfig.from_dict({"world_size": 28.0, "num_drones": 8, "bogus": 1, "seed": 3})
    assert c.world_size == 28.0 and c.seed == 3 and c.max_steps == 400 and c.neighbor_k == 3
    assert swarm_b200.DroneEnvConfig.from_dict(None) == swarm_b200.DroneEnvConfig()


def test_unpack_flags_is_the_inverse_of_the_packing_rule():
    """SwarmHostOut.flags (ABI 5): bit k of the byte is FLAG_FIELDS[k]; `SwarmEngine.unpack_flags` splits it (host side)."""
    import numpy as np
    import swarm_b200
    from swarm_b200 import _abi
    rng = np.random.default_rng(0)
    fields = {n: rng.integers(0, 2, size=(7, 5)).astype(np.uint8) for n in _abi.FLAG_FIELDS}
    packed = sum(fields[n] << k for k, n in enumerate(_abi.FLAG_FIELDS)).astype(np.uint8)
    assert packed.max() < 32
    un = swarm_b200.SwarmEngine.unpack_flags(packed)
    assert set(un) == set(_abi.FLAG_FIELDS)
    for n in _abi.FLAG_FIELDS:
        assert np.array_equal(un[n], fields[n].astype(bool)), n
    assert (_abi.FLAG_TERMINATED, _abi.FLAG_TRUNCATED, _abi.FLAG_REACHED, _abi.FLAG_COLLISION, _abi.FLAG_OBS_VALID) == (1, 2, 4, 8, 16)
    assert set(swarm_b200.SwarmEngine.OUTPUT_SETS) == {"packed", "lean"}
    assert "flags" in swarm_b200.SwarmEngine.OUTPUT_SETS["lean"] and "global_state" not in swarm_b200.SwarmEngine.OUTPUT_SETS["lean"]
