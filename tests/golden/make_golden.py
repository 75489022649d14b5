"""Generates the committed golden fixtures.

  python tests/golden/make_golden.py            # CPU: Oracle-I / Oracle-L outputs (oracle_*.npz)
  python tests/golden/make_golden.py reference  # GPU box: outputs of the UNMODIFIED reference
                                                # (oracle/_ref/libref_step.so) -> gpurun_out/, copied here
  python tests/golden/make_golden.py reference_ic  # CPU: digests of the reference's own initial conditions
                                                # (its main() run up to the uploads, oracle/ref_wrap.cu:ref_ic)

The reference ships no fixtures; reference_b200_*.npz are outputs of its own kernels +
simulationStep() run on a B200, so the CPU-only suite can pin Oracle-L without a GPU.
"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nbody_barnes_hut_cuda_b200 as bh  # noqa: E402
import oracle_lib as O  # noqa: E402

N_REF = 8192
N_ORC = 4096


def inputs(n):
    return bh.ic_uniform_cube(n, 42, 1000.0)


def oracle_golden():
    soa = inputs(N_ORC)
    posm, vel, ids = O.soa_to_internal(soa)
    one = O.engine_step(posm, vel, ids, 1)
    ten = O.engine_step(posm, vel, ids, 10)
    b = O.bounds(*soa[:3])
    keys, idx = O.morton_keys(*soa[:3], b)
    ks, perm = O.stable_sort(keys, idx)
    meta, child, root = O.tree_build(ks)
    tuples = sorted(O.cell_tuples(meta, ks))
    lit = O.reference_step(soa, 1, fixed=0)
    np.savez_compressed(
        os.path.join(HERE, f"oracle_uniform{N_ORC}.npz"), n=N_ORC, seed=42, bounds=b, keys=keys, perm=perm,
        tree_tuples=np.array(tuples, np.int64), root=root, acc_step1=one["acc"][:, :3], ids_step1=one["ids"],
        inter_step1=np.array([one["inter_cell"], one["inter_body"], one["cells"]], np.int64),
        posm_step10=ten["posm"], vel_step10=ten["vel"][:, :3], ids_step10=ten["ids"],
        literal_acc=np.stack([lit["ax"], lit["ay"], lit["az"]], 1), literal_nodes=lit["nodes"])
    print("wrote oracle golden", N_ORC)


IC_SIZES = (1000, 16384, 500_000, 1_000_000)   # 500,000 = the reference default (bench:31), 1,000,000 = configs[1]


def reference_ic(n):
    """The seven host arrays the unmodified reference uploads (bench:294-308, 329-335)."""
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_step.so"))
    a = [np.zeros(n, np.float32) for _ in range(7)]
    assert L.ref_ic(n, *[x.ctypes.data_as(C.c_void_p) for x in a]) == 0
    return a


def ic_digest(arrays):
    return hashlib.sha256(b"".join(np.ascontiguousarray(x, np.float32).tobytes() for x in arrays)).hexdigest()


def reference_ic_golden():
    import json

    out = {str(n): ic_digest(reference_ic(n)) for n in IC_SIZES}
    head = {str(n): [float(x[0]) for x in reference_ic(n)] for n in IC_SIZES[:1]}
    with open(os.path.join(HERE, "reference_ic_sha256.json"), "w") as f:
        json.dump({"what": "sha256 over posX|posY|posZ|velX|velY|velZ|mass (float32) as produced by the unmodified "
                           "reference main(), nbody_v5_bench.cu:294-308, srand(42)", "sha256": out, "first_body_n1000": head["1000"]}, f, indent=1)
    print("wrote reference_ic_sha256.json", out)


def reference_golden():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_reference_pin import run_reference

    soa = inputs(N_REF)
    r1 = run_reference(soa, 1)
    r3 = run_reference(soa, 3)
    out = os.path.join(ROOT, "gpurun_out", f"reference_b200_uniform{N_REF}.npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, n=N_REF, seed=42, bounds=r1["bounds"], sorted_keys=r1["keys"], sorted_idx=r1["idx"],
                        acc=np.stack([r1["ax"], r1["ay"], r1["az"]], 1),
                        pos1=np.stack([r1["px"], r1["py"], r1["pz"]], 1), vel1=np.stack([r1["vx"], r1["vy"], r1["vz"]], 1),
                        pos3=np.stack([r3["px"], r3["py"], r3["pz"]], 1), nodes=r1["nodes"],
                        sorted_keys3=r3["keys"], sorted_idx3=r3["idx"])
    print("wrote", out)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "reference_ic":
    reference_ic_golden()
    sys.exit(0)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "reference":
        reference_golden()
    else:
        oracle_golden()
