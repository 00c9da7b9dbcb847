"""CPU-side checks of the drop-in links (INTEGRATION.md sections 2 and 5), run where there is no GPU:

  * both GPU-linked Moldy programs take the hot path from libmoldy_b200.so and fail LOUDLY without a CUDA device
    (there is no CPU path), instead of computing anything on the host;
  * in moldy_gpu_evalf the program's `eval_forces` is the trampoline (a jump to mdb_eval_forces_moldy), and the
    library's mdb_eval_forces_moldy does not call back through the exported `eval_forces` name -- which inside the
    program binds to that trampoline and would loop for ever (the bug of profiles/r01_s4_evalf.md)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
LIB = os.path.join(ROOT, "moldy_b200", "libmoldy_b200.so")


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="a CUDA device is present: the programs would simply run")
@pytest.mark.parametrize("prog", ["moldy_gpu", "moldy_gpu_evalf"])
def test_gpu_linked_moldy_fails_loudly_without_a_device(prog, tmp_path):
    binary = os.path.join(REFDIR, prog)
    if not os.path.exists(binary):
        pytest.skip(f"oracle/_ref/{prog} not built (make -C oracle ref)")
    from tests import test_gpu_dropin as t
    shutil.copy(os.path.join(ROOT, "tests", "golden", "tip4p_256_eq.txt"), tmp_path)
    (tmp_path / "control").write_text(t.CONTROL.format(nsteps=2, every=1, rdf=0, rdfout=1000000))
    out = subprocess.run([binary, "control"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert "*F* libmoldy_b200" in out.stdout and "no CPU path" in out.stdout, out.stdout[-1500:]
    assert "Timestep 1" not in out.stdout


def _disasm(path, symbol):
    txt = subprocess.run(["objdump", "-d", "--no-show-raw-insn", path], capture_output=True, text=True).stdout
    body, on = [], False
    for ln in txt.splitlines():
        if ln.endswith(f"<{symbol}>:"):
            on = True
            continue
        if on:
            if not ln.strip():
                break
            body.append(ln)
    return "\n".join(body)


@pytest.mark.skipif(shutil.which("objdump") is None, reason="binutils not installed")
def test_eval_forces_trampoline_and_no_interposable_self_call():
    prog = os.path.join(REFDIR, "moldy_gpu_evalf")
    if not os.path.exists(prog):
        pytest.skip("oracle/_ref/moldy_gpu_evalf not built (make -C oracle ref)")
    tramp = _disasm(prog, "eval_forces")
    assert "mdb_eval_forces_moldy" in tramp and len(tramp.splitlines()) <= 4, tramp
    inner = _disasm(LIB, "mdb_eval_forces_moldy")
    assert inner and "<eval_forces" not in inner and "eval_forces@plt" not in inner, inner
    assert "eval_forces_impl" in inner
