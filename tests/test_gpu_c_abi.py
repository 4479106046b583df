"""The C ABI is language-neutral: a plain-C program (tests/c_abi/smoke.c, gcc + libcudart, no Python / torch / C++ types)
drives dpa_forward / dpa_backward on a batch from a file and reproduces the Python binding's results bit for bit."""
import os
import re
import struct
import subprocess

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_caller_matches_python_binding(tmp_path):
    from halva_b200 import _lib, dpa_loss, read_status
    from halva_b200.synth import CONFIGS, make_inputs
    exe = tmp_path / "smoke"
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    subprocess.run(["gcc", "-O1", os.path.join(ROOT, "tests", "c_abi", "smoke.c"), f"-I{ROOT}/include", f"-I{cuda}/include",
                    f"-L{cuda}/lib64", "-lcudart", "-ldl", "-lm", "-o", str(exe)], check=True)
    shape = CONFIGS["small"]
    t = make_inputs(shape, seed=13, device="cuda")
    d_loss = 0.5
    raw = lambda x: (x.view(torch.int16) if x.dtype == torch.bfloat16 else x).contiguous().cpu().numpy().tobytes()
    batch = tmp_path / "batch.bin"
    with open(batch, "wb") as f:
        f.write(struct.pack("<8i", shape.B, shape.T, shape.B_ref, shape.T_ref, shape.H, shape.V, 0, 0))
        f.write(struct.pack("<2f", shape.alpha, d_loss))
        for k in ("h_pair", "labels_pair", "signs_pair", "h_ref_pol", "h_ref_frz", "labels_ref", "W"):
            f.write(raw(t[k]))
    env = dict(os.environ, LD_LIBRARY_PATH=f"{cuda}/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([str(exe), _lib.LIB_PATH, str(batch)], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-1500:]
    num = lambda line: [float(x) for x in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?|nan|inf", line)]
    lines = out.stdout.strip().splitlines()
    st_c, loss_c, grads_c = num(lines[0]), num(lines[1]), num(lines[2])
    hp = t["h_pair"].clone().requires_grad_(True)
    hr = t["h_ref_pol"].clone().requires_grad_(True)
    W = t["W"].clone().requires_grad_(True)
    loss, align, div = dpa_loss(hp, t["labels_pair"], t["signs_pair"], hr, t["labels_ref"], W, h_ref_frozen=t["h_ref_frz"], alpha=shape.alpha)
    (loss * d_loss).backward()
    st = read_status(loss)
    assert st_c == [st["status"], st["K"], st["n_pair_active"], st["n_ref_active"]]
    f32 = lambda x: struct.unpack("<f", struct.pack("<f", float(x)))[0]
    assert [f32(v) for v in loss_c] == [loss.item(), align.item(), div.item()]            # the same kernels: bit-identical
    for got, g in zip(grads_c, (hp.grad, hr.grad, W.grad)):
        want = g.float().abs().double().sum().item()
        assert abs(got - want) <= 1e-9 * max(want, 1.0), (got, want)
