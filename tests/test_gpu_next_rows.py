"""Callers either side of the step path: visual export, diagnostics, state I/O, bench front-end."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
f = np.float32


def test_visual_export_matches_updateVisualsKernel(bh):
    """nbody_v5.cu:278-292: interleaved xyz + speed colour ramp, original body order."""
    import torch

    n = 5000
    soa = bh.ic_refdisk(n, 42)
    with bh.BHEngine(n) as eng:
        eng.load_soa(*soa)
        eng.simulation_step(2)
        vp = torch.zeros(3 * n, dtype=torch.float32, device="cuda")
        vc = torch.zeros(3 * n, dtype=torch.float32, device="cuda")
        eng.export_visuals(vp, vc)
        torch.cuda.synchronize()
        px, py, pz, vx, vy, vz = eng.read_soa(want_acc=False)
    p = vp.cpu().numpy().reshape(n, 3)
    c = vc.cpu().numpy().reshape(n, 3)
    assert p[:, 0].tobytes() == px.tobytes() and p[:, 1].tobytes() == py.tobytes() and p[:, 2].tobytes() == pz.tobytes()
    speed = np.sqrt(vx.astype(np.float64) ** 2 + vy.astype(np.float64) ** 2 + vz.astype(np.float64) ** 2)
    t = np.minimum(speed / 150.0, 1.0)
    want = np.stack([0.4 + 0.6 * t, 0.3 + 0.4 * t, 1.0 - 0.7 * t], 1)
    assert np.abs(c - want).max() < 1e-6


def test_momentum_and_its_conservation(bh):
    n = 20000
    soa = bh.ic_plummer(n, 3, 200.0, 10.0, 4.5, 0.5)
    m = soa[6].astype(np.float64)
    pos = np.stack(soa[:3], 1).astype(np.float64)
    vel = np.stack(soa[3:6], 1).astype(np.float64)
    want = np.concatenate([[m.sum()], (m[:, None] * vel).sum(0), np.cross(pos, m[:, None] * vel).sum(0)])
    with bh.BHEngine(n) as eng:
        eng.load_soa(*soa)
        got = eng.momentum()
        assert np.allclose(got, want, rtol=1e-9, atol=1e-6)
        eng.simulation_step(20)
        after = eng.momentum()
    scale = (m[:, None] * np.abs(vel)).sum()
    # Barnes-Hut forces are not exactly antisymmetric: momentum drifts only at the multipole-error level
    assert np.abs(after[1:4] - got[1:4]).max() < 1e-4 * scale


def test_text_dump_has_the_old_tools_format(bh, tmp_path):
    """output_bh.txt:1-5."""
    n = 300
    soa = bh.ic_uniform_cube(n, 9, 1000.0)
    path = str(tmp_path / "out.txt")
    with bh.BHEngine(n) as eng:
        eng.load_soa(*soa)
        eng.simulation_step(4)
        eng.dump_text(path)
        ref = eng.read_soa(want_acc=False)
    lines = open(path).read().splitlines()
    assert lines[0] == "# Barnes-Hut N-Body Simulation Results"
    assert lines[1] == "# Final positions and velocities after 4 steps"
    assert lines[2] == "# Bodies: 300, Theta: 0.50, dt: 0.020" and lines[3] == "# Format: x y z vx vy vz"
    rows = np.array([[float(x) for x in ln.split()] for ln in lines[4:]])
    assert rows.shape == (n, 6) and np.abs(rows - np.stack(ref, 1)).max() < 1e-6 * 1000 + 1e-6


def test_checkpoint_resume_is_bit_exact(bh, tmp_path):
    n = 30000
    soa = bh.ic_refdisk(n, 42)
    path = str(tmp_path / "ckpt.bin")
    with bh.BHEngine(n) as eng:
        eng.load_soa(*soa)
        eng.simulation_step(3)
        eng.save_checkpoint(path)
        eng.simulation_step(4)
        want = (eng.debug_get(bh.DBG.POSM), eng.debug_get(bh.DBG.VEL), eng.debug_get(bh.DBG.IDS))
    with bh.BHEngine(n) as eng:
        eng.load_checkpoint(path)
        assert eng.stat(bh.STAT.STEPS) == 3 and eng.n == n
        eng.simulation_step(4)
        got = (eng.debug_get(bh.DBG.POSM), eng.debug_get(bh.DBG.VEL), eng.debug_get(bh.DBG.IDS))
    assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes() and (got[2] == want[2]).all()
    with bh.BHEngine(100) as small:
        with pytest.raises(bh.BHError):
            small.load_checkpoint(path)          # does not fit: refused, not truncated


def test_checkpoint_loader_validates_what_bh_create_validates(bh, tmp_path):
    """parameters from a file get the range checks of bh_create; a different key width is refused."""
    import struct

    n = 2000
    soa = bh.ic_refdisk(n, 42)
    path = str(tmp_path / "ckpt.bin")
    with bh.BHEngine(n) as eng:
        eng.load_soa(*soa)
        eng.simulation_step(1)
        eng.save_checkpoint(path)
    raw = bytearray(open(path, "rb").read())
    magic, version, hbytes, nn, steps, theta, G, dt, soft, vmax, key_bits, leaf_cap, _ = struct.unpack_from("<8siiqqfffffiii", raw, 0)
    assert magic == b"BHB200\0\0" and version == 2 and nn == n and steps == 1 and key_bits == 30 and leaf_cap == 1
    assert (theta, G, dt, soft, vmax) == (0.5, 0.5, struct.unpack("<f", struct.pack("<f", 0.02))[0], 50.0, 500.0)
    bad = str(tmp_path / "bad.bin")
    off_soft = 8 + 4 + 4 + 8 + 8 + 3 * 4
    for patch in (struct.pack("<f", 0.0), struct.pack("<f", -1.0), struct.pack("<f", float("nan"))):
        b2 = bytearray(raw)
        b2[off_soft:off_soft + 4] = patch
        open(bad, "wb").write(b2)
        with bh.BHEngine(n) as eng, pytest.raises(bh.BHError):
            eng.load_checkpoint(bad)                     # softening <= 0 would make the self term NaN
    with bh.BHEngine(n, key_bits=60) as eng, pytest.raises(bh.BHError):
        eng.load_checkpoint(path)                        # 30-bit file, 60-bit context: a different tree
    b2 = bytearray(raw)
    b2[8:12] = struct.pack("<i", 1)                      # the round-1 layout embedded the ABI struct: refused
    open(bad, "wb").write(b2)
    with bh.BHEngine(n) as eng, pytest.raises(bh.BHError):
        eng.load_checkpoint(bad)


def test_bench_front_end_prints_the_reference_table(bh):
    """nbody_v5_bench.cu:287,350-351,366."""
    exe = os.path.join(os.path.dirname(bh.library_path()), "nbody_bench")
    if not os.path.exists(exe):
        pytest.skip("front-end not built")
    r = subprocess.run([exe, "--n", "20000", "--frames", "3", "--phases"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = r.stdout.splitlines()
    assert out[0] == "Pokretanje Benchmarka za N = 20000..."
    assert any(ln.startswith("Frame      | Trajanje (ms)   | FPS") for ln in out)
    assert sum(1 for ln in out if ln[:1].isdigit() and "|" in ln) == 3
    assert any("interactions/s" in ln for ln in out) and any("octree build" in ln for ln in out)


def test_the_reference_bench_itself_runs_on_libbh(bh):
    """INTEGRATION.md, compiled: /root/reference/nbody_v5_bench.cu with the documented patch applied by
    tools/integration/patch_reference.py (simulationStep() -> bh_step, bh_create/bh_import_soa after the reference's
    own uploads, bh_export_soa into its arrays before its cudaFrees) and linked with -lbh.  Its own main() generates
    the disk, prints its own table; the state it reads back equals the engine driven through the Python binding."""
    import re

    exe = os.path.join(os.path.dirname(os.path.dirname(bh.library_path())), "oracle", "_ref", "nbody_v5_bench_libbh")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/nbody_v5_bench_libbh not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout
    assert "Pokretanje Benchmarka za N = 500000" in out                      # bench:287 with the reference's default N (bench:31)
    frames = re.findall(r"^(\d+)\s+\|\s+([0-9.]+)\s+\|\s+([0-9.]+)", out, re.M)
    assert len(frames) == 100 and [int(f[0]) for f in frames] == list(range(100))
    m = re.search(r"libbh: sum\(posX\) after 100 frames = ([-0-9.eE+]+), interactions/body = ([0-9.]+), device error flag = (\d+)", out)
    assert m and int(m.group(3)) == 0 and 500 < float(m.group(2)) < 3000
    n = 500_000
    soa = bh.ic_refdisk(n, 42)
    with bh.BHEngine(n) as eng:
        eng.load_soa(*soa)
        eng.simulation_step(100)
        want = eng.read_soa(want_acc=False)[0].astype(np.float64).sum()
    assert abs(float(m.group(1)) - want) <= 1e-6 * max(1.0, abs(want))
    assert np.median([float(f[1]) for f in frames]) < 5.0                    # ms per frame at 500k bodies (reference: ~30)
