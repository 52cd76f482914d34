#!/usr/bin/env python3
"""oracle/build_ref.py -- TEST INFRASTRUCTURE. Builds oracle/_ref/kmcref_<tag>[,_shipped] from the
UNMODIFIED reference translation unit, compiled from where it lies (/root/reference/main.cpp).

The default system (tag n200) compiles the file as it is. Other sizes need different values for the
ten size #defines (main.cpp:47-69, compile-time array extents); for those a scratch copy with ONLY
those ten lines rewritten is generated under /tmp (never inside this repository) and compiled.
Outputs go only to oracle/_ref/ (git-ignored; it does travel to the GPU box with gpurun).

Recipe:  g++ -O2 -fPIC -ffp-contract=off -c main.cpp
                     objcopy --redefine-sym main=ref_main [--weaken-symbol=_Z5rand2v]
                     g++ ref_harness.cpp main_*.o
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("KMC_REFERENCE", "/root/reference/main.cpp")
TMP = os.environ.get("KMC_ORACLE_TMP", "/tmp/kmc_oracle_build")
VARIANTS = {"n200": (150, 50), "n2000": (1500, 500), "n40": (30, 10), "n20000": (15000, 5000)}
DEFAULT_TAGS = ["n200", "n2000", "n40"]


def sh(*cmd):
    subprocess.run(list(cmd), check=True)


def build(tag):
    na, nb = VARIANTS[tag]
    os.makedirs(TMP, exist_ok=True)
    os.makedirs(os.path.join(HERE, "_ref"), exist_ok=True)
    src = REF
    if tag != "n200":
        text = open(REF).read()
        defs = {"RB_A_tot_num": na * 4, "protein_A_tot_num": na, "protein_A_tot_num_matrix": na + 1,
                "RB_B_tot_num": nb * 4, "protein_B_tot_num": nb, "protein_B_tot_num_matrix": nb + 1,
                "protein_tot_num": na + nb, "max_bond_num": 6 * na, "protein_tot_num_matrix": na + nb + 1,
                "max_bond_num_matrix": 6 * na + 1}
        for name, val in defs.items():
            text, cnt = re.subn(r"(?m)^#define %s\s+\d+" % name, "#define %s %d" % (name, val), text)
            assert cnt == 1, (name, cnt)
        src = os.path.join(TMP, "main_%s.cpp" % tag)
        open(src, "w").write(text)
    obj = os.path.join(TMP, "main_%s.o" % tag)
    flags = ["-O2", "-fPIC", "-g", "-w", "-ffp-contract=off"] + (["-mcmodel=medium"] if na + nb >= 10000 else [])
    sh("g++", *flags, "-c", src, "-o", obj)
    inj, ship = obj.replace(".o", "_inj.o"), obj.replace(".o", "_ship.o")
    sh("objcopy", "--redefine-sym", "main=ref_main", "--weaken-symbol=_Z5rand2v", obj, inj)
    # call sites of rand2() inside ref_main -> source lines (DWARF), so that the injected rand2 can key its draw by call site:
    # return address - &ref_main  ->  main.cpp line
    dis = subprocess.run(["objdump", "-dr", "--no-show-raw-insn", inj], capture_output=True, text=True, check=True).stdout.splitlines()
    sites = []
    for n, line in enumerate(dis):
        if "R_X86_64_PLT32" in line and "_Z5rand2v" in line:
            addr = int(dis[n - 1].split(":")[0].strip(), 16)
            src_line = subprocess.run(["addr2line", "-e", inj, "-j", ".text.startup", hex(addr)], capture_output=True, text=True, check=True).stdout
            sites.append((addr + 5, int(re.search(r":(\d+)", src_line).group(1))))
    assert len(sites) == 30, len(sites)
    with open(os.path.join(TMP, "callsites_%s.h" % tag), "w") as f:
        f.write("static const struct { long ret; int line; } g_sites[] = {%s};\n" % ", ".join("{%d, %d}" % s for s in sites))
    sh("objcopy", "--redefine-sym", "main=ref_main", obj, ship)
    hflags = ["-O2", "-DKMC_NA=%d" % na, "-DKMC_NB=%d" % nb, "-I", TMP, "-I", HERE, "-DKMC_CALLSITES=\"callsites_%s.h\"" % tag] + (["-mcmodel=medium"] if na + nb >= 10000 else [])
    harness = os.path.join(HERE, "ref_harness.cpp")
    sh("g++", *hflags, harness, inj, "-o", os.path.join(HERE, "_ref", "kmcref_" + tag))
    sh("g++", *hflags, "-DKMC_SHIPPED_RAND2", harness, ship, "-o", os.path.join(HERE, "_ref", "kmcref_%s_shipped" % tag))


if __name__ == "__main__":
    if not os.path.exists(REF):
        print("reference source %s not present: keeping prebuilt oracle/_ref as is" % REF)
        sys.exit(0)
    for t in (sys.argv[1:] or DEFAULT_TAGS):
        build(t)
        print("built oracle/_ref/kmcref_%s" % t)
