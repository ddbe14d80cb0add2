"""The product's box-box geometry (genesis_b200/csrc/gsr_boxbox.cuh bb_contacts, RigidOptions.box_box_detection) compiled for the
HOST and run next to the oracle's restatement of the reference's func_box_box_contact (oracle/box_box.hpp, itself pinned to the
reference by tests/golden/ref_boxbox_*.npz) on 20000 random box pairs: same number of contacts, positions / normals / penetrations
within 2e-5.  CPU tier: no GPU needed (the GPU tier runs the same routine inside k_step, tests/test_gpu_parity.py)."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def test_product_box_box_geometry_matches_oracle_on_host():
    exe = os.path.join(tempfile.mkdtemp(), "boxbox_host")
    src = os.path.join(HERE, "host_harness", "boxbox_host.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-I/usr/local/cuda/include", "-o", exe, src], check=True)
    for seed in (1, 2):
        r = subprocess.run([exe, "10000", str(seed)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        words = r.stdout.split()
        assert int(words[3]) > 8000 and int(words[5]) == 0, r.stdout  # most pairs collide, none disagrees
