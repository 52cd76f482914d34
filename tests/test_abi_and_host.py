"""CPU: the C-ABI library loads, exports every symbol include/kmc_b200.h declares, fails loudly without a GPU, and its
host-side logic (record formatting, parameter defaults, Philox stream, rank sharding) is right. No device compute here."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import kmc_b200
import pyoracle
from common import apply_regime

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_kat.json")))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "kmc_b200.h")).read()
    declared = set(re.findall(r"\b(kmc_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"kmc_status"}
    assert len(declared) >= 25
    lib = kmc_b200.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), "libkmc_b200.so does not export " + name
    assert declared == set(kmc_b200.EXPORTS), declared ^ set(kmc_b200.EXPORTS)
    assert lib.kmc_abi_version() == 1


def test_default_params_are_the_reference_globals():
    """main.cpp:39-99 as shipped"""
    p = kmc_b200.default_params()
    assert (p.box[0], p.box[1], p.box[2], p.dt, p.pai) == (5773, 5773, 1000, 10, 3.1415926)
    assert (p.rA, p.DA, p.DrotA, p.rB, p.DB, p.DrotB) == (20, 1, 0.0174, 30, 7.2614, 0.0061209)
    assert (p.mono_cis_on, p.mono_cis_off, p.cis_D, p.cis_Drot, p.cis_on, p.cis_off) == (0.000047, 0.000000000000112, 0.5, 0.005, 0.00096, 0.000000000000112)
    assert (p.bond_D, p.bond_Drot, p.on, p.off) == (0.5, 0.005, 0.04, 0.000000000000348)
    assert (p.bond_dist_cut, p.thetapd_cut, p.thetaot_cut, p.cis_thetaot_cut, p.cis_dist_cut) == (18, 45, 90, 10, 15)
    assert (p.n_receptor, p.n_ligand, p.n_replicas) == (150, 50, 1)
    # the oracle's defaults are the same numbers (two independent transcriptions of main.cpp:39-99)
    q = pyoracle.default_params()
    for f, _ in pyoracle.Params._fields_:
        if hasattr(p, f) and f not in ("box", "seed"):
            assert getattr(p, f) == getattr(q, f), f


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_cpu_fallback():
    """the product path must fail loudly when there is no device (no CPU fallback, no oracle behind it)"""
    with pytest.raises(kmc_b200.KmcError, match="no CUDA device|CUDA"):
        kmc_b200.Kmc(kmc_b200.default_params())
    src = open(os.path.join(ROOT, "kmc-with-a-diffusion-reaction-algorithm_b200", "csrc", "kmc_engine.cu")).read()
    src += open(os.path.join(ROOT, "kmc-with-a-diffusion-reaction-algorithm_b200", "kmc_b200", "__init__.py")).read()
    assert "oracle" not in src.replace("no oracle", "")


def test_records_byte_compatible_with_reference():
    """bond.dat lines and cluster.log frames formatted by the library == the text the unmodified reference wrote
    (tests/golden/ref_kat.json: ref_records), fed with the oracle's numbers at the same steps."""
    p = apply_regime(pyoracle.default_params(box=(2500, 2500, 400), use_grid=1), "dense")
    o = pyoracle.Oracle(p)
    o.init_reference()
    bond, cluster = "", ""
    for step in (5000, 10000):
        o.step(5000)
        c = o.counts()
        bond += kmc_b200.format_bond_dat(p.dt, step, c["bond_num_rl"], c["bond_num_mono_cis"], c["bond_num_cis"], c["bond_num"],
                                         c["cluster_size"], c["max_complex"])
        cluster += kmc_b200.format_cluster_log(p.dt, step, o.results())
    assert bond == KAT["ref_records"]["bond_dat"]
    assert cluster == KAT["ref_records"]["cluster_log"]


def test_cluster_log_large_time_uses_default_float_format():
    # main.cpp:2293 prints mc_time_step*time_step with the default ostream format: 1e6 -> "1e+06"
    assert kmc_b200.format_cluster_log(10.0, 100000, [[151]]).startswith("Hello Cluster!, t=1e+06\n151  \n")
    assert kmc_b200.format_bond_dat(10.0, 5000, 1, 0, 0, 1, 2.0, 2) == "      50000.000    1    0         0         1     2.000         2\n"


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10, through the oracle's implementation (the device implementation is
    compared with it draw by draw in the GPU replay tests)."""
    code = r'''
#include "philox.h"
#include <cstdio>
int main(){ auto a=kmco::philox4x32_10(0,0,0,0,0,0); auto b=kmco::philox4x32_10(0xffffffffu,0xffffffffu,0xffffffffu,0xffffffffu,0xffffffffu,0xffffffffu);
 auto c=kmco::philox4x32_10(0x243f6a88u,0x85a308d3u,0x13198a2eu,0x03707344u,0xa4093822u,0x299f31d0u);
 printf("%08x %08x %08x %08x\n%08x %08x %08x %08x\n%08x %08x %08x %08x\n",a.v[0],a.v[1],a.v[2],a.v[3],b.v[0],b.v[1],b.v[2],b.v[3],c.v[0],c.v[1],c.v[2],c.v[3]);
 printf("%.17g %d\n", kmco::keyed_uniform(1,2,3,4,5), kmco::keyed_rand31(1,2,3,4)); }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.cpp"), "w").write(code)
        subprocess.run(["g++", "-O1", "-I", os.path.join(ROOT, "oracle"), os.path.join(td, "t.cpp"), "-o", os.path.join(td, "t")], check=True)
        out = subprocess.run([os.path.join(td, "t")], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0] == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"
    assert out[1] == "408f276d 41c83b0e a20bc7c6 6d5451fd"
    assert out[2] == "d16cfe09 94fdcceb 5001e420 24126ea1"
    u, r = out[3].split()
    assert 0.0 <= float(u) < 1.0 and 0 <= int(r) < 2 ** 31


def test_replica_partition():
    from kmc_b200.sharding import replica_range, rank_seed
    for world in (1, 2, 3, 4, 8):
        for n in (1, 7, 64, 1024):
            cover = []
            for r in range(world):
                lo, hi = replica_range(r, world, n)
                cover += list(range(lo, hi))
                assert rank_seed(100, r, world, n) == 100 + lo
            assert cover == list(range(n))


def test_every_entry_point_is_documented_in_integration_md():
    """INTEGRATION.md maps each C-ABI entry point to the reference interface it replaces: no declared function may be missing there"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "kmc_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    declared = set(re.findall(r"\b(kmc_[a-z_0-9]+)\s*\(", hdr)) - {"kmc_status"}
    families = [m[:-1] for m in re.findall(r"kmc_[a-z_]+\*", doc)]          # e.g. kmc_strip_*, kmc_profile*
    missing = sorted(d for d in declared if d not in doc and not any(d.startswith(f) for f in families))
    assert not missing, missing


def test_alignment_windows_equal_the_square_root_test():
    """The kernels replace AreSame(sqrt(q), D) (main.cpp:2368-2371 on the alignment distances of 1205-1215 / 1245-1255) by a window
    on the squared distance q. Host arithmetic only: for the four alignment lengths of the default parameters (and a few odd ones)
    the window must agree with the square-root test on every double within 4096 ulps of either end, and on a coarse sweep beyond."""
    import numpy as np
    import kmc_b200
    p = kmc_b200.default_params()
    lengths = [p.bond_dist_cut / 2 + p.rA + p.rB, p.bond_dist_cut / 2, p.cis_dist_cut / 2 + p.rA + p.rA, p.cis_dist_cut / 2, 1.0, 0.37, 123.456, 1e4 / 3]
    for D in lengths:
        w = kmc_b200.alignment_window(D)
        assert w is not None and w[0] < D * D < w[1]
        for edge in w:
            q = np.full(8193, edge)
            for i in range(4096):                                   # neighbouring doubles on both sides of the edge
                q[4095 - i] = np.nextafter(q[4096 - i], 0.0)
                q[4097 + i] = np.nextafter(q[4096 + i], np.inf)
            ref = np.abs(np.sqrt(q) - D) < 1.0e-8                    # numpy's float64 sqrt is correctly rounded, like C's and CUDA's
            got = (q >= w[0]) & (q <= w[1])
            assert np.array_equal(ref, got), D
        q = np.linspace(0.0, 4 * D * D, 20001)
        assert np.array_equal(np.abs(np.sqrt(q) - D) < 1.0e-8, (q >= w[0]) & (q <= w[1]))
    assert kmc_b200.alignment_window(0.0) is None                    # degenerate length: the kernels keep the square root
