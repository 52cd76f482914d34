"""oracle/refio.py -- TEST INFRASTRUCTURE. Readers/writers for the frame files exchanged with
oracle/_ref/kmcref_* (the unmodified reference driven by oracle/ref_harness.cpp) and helpers to run
it. Frame layout (little endian): int64 step | int32 cnt[8] = bond_num, bond_num_rl, bond_num_cis,
bond_num_mono_cis, max_complex, tot_cluster_num, tot_proteins_in_cluster, N | float64 cluster_size |
float64 R_x,R_y,R_z [N+1][5][5] | int32 protein_status[N+1][5] | int32 res_nei[N+1][7]
(the reference's own array shapes, main.cpp:102-118)."""
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def frame_nbytes(n):
    return 8 + 32 + 8 + 3 * (n + 1) * 25 * 8 + (n + 1) * 5 * 4 + (n + 1) * 7 * 4


def read_frames(path, n):
    """Returns a list of dicts, one per frame."""
    raw = np.fromfile(path, dtype=np.uint8)
    fb = frame_nbytes(n)
    assert raw.size % fb == 0, (raw.size, fb)
    out = []
    for f in range(raw.size // fb):
        b = raw[f * fb:(f + 1) * fb]
        o = 0
        step = int(b[o:o + 8].view(np.int64)[0]); o += 8
        cnt = b[o:o + 32].view(np.int32).copy(); o += 32
        cs = float(b[o:o + 8].view(np.float64)[0]); o += 8
        m = (n + 1) * 25 * 8
        R = np.stack([b[o + c * m:o + (c + 1) * m].view(np.float64).reshape(n + 1, 5, 5) for c in range(3)], axis=-1)
        o += 3 * m
        st = b[o:o + (n + 1) * 20].view(np.int32).reshape(n + 1, 5).copy(); o += (n + 1) * 20
        rn = b[o:o + (n + 1) * 28].view(np.int32).reshape(n + 1, 7).copy(); o += (n + 1) * 28
        assert cnt[7] == n
        out.append(dict(step=step, bond_num=int(cnt[0]), bond_num_rl=int(cnt[1]), bond_num_cis=int(cnt[2]),
                        bond_num_mono_cis=int(cnt[3]), max_complex=int(cnt[4]), tot_cluster_num=int(cnt[5]),
                        tot_proteins_in_cluster=int(cnt[6]), cluster_size=cs, R=R.copy(), status=st, res_nei=rn))
    return out


def write_frame(path, fr):
    n = fr["R"].shape[0] - 1
    with open(path, "wb") as f:
        f.write(np.int64(fr["step"]).tobytes())
        cnt = np.array([fr["bond_num"], fr["bond_num_rl"], fr["bond_num_cis"], fr["bond_num_mono_cis"],
                        fr["max_complex"], 0, 0, n], dtype=np.int32)
        f.write(cnt.tobytes())
        f.write(np.float64(fr.get("cluster_size", 0.0)).tobytes())
        for c in range(3):
            f.write(np.ascontiguousarray(fr["R"][..., c], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(fr["status"], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(fr["res_nei"], dtype=np.int32).tobytes())


def fnv1a64(*arrays):
    h = 1469598103934665603
    for a in arrays:
        for byte in np.ascontiguousarray(a).view(np.uint8).ravel().tolist():
            h = ((h ^ byte) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def ref_available(tag="n200"):
    return os.path.exists(os.path.join(REF_DIR, "kmcref_" + tag))


def run_ref(tag, n, steps, sets=None, scales=None, rand2_state=None, rand_state=None, frames_every=0,
            in_frame=None, shipped=False, timeout=3600, keyed_seed=None):
    """Runs oracle/_ref/kmcref_<tag>; returns (summary dict, list of frames)."""
    exe = os.path.join(REF_DIR, "kmcref_" + tag + ("_shipped" if shipped else ""))
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "frames.bin")
        cmd = [exe, "--steps", str(steps), "--out", out, "--workdir", os.path.join(td, "wd")]
        for k, v in (sets or {}).items():
            cmd += ["--set", "%s=%r" % (k, float(v))]
        for k, v in (scales or {}).items():
            cmd += ["--scale", "%s=%r" % (k, float(v))]
        if rand2_state is not None:
            cmd += ["--rand2-state", str(rand2_state)]
        if rand_state is not None:
            cmd += ["--rand-state", str(rand_state)]
        if frames_every:
            cmd += ["--frames-every", str(frames_every)]
        if keyed_seed is not None:
            cmd += ["--keyed", str(keyed_seed)]
        if in_frame is not None:
            inp = os.path.join(td, "in.bin")
            write_frame(inp, in_frame)
            cmd += ["--in", inp]
        p = subprocess.run(cmd, cwd=td, capture_output=True, text=True, timeout=timeout)
        if p.returncode != 0:
            raise RuntimeError("kmcref failed: %s\n%s" % (p.stdout, p.stderr))
        summary = json.loads(p.stdout.strip().splitlines()[-1])
        frames = read_frames(out, n)
    return summary, frames
