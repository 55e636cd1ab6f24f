"""Static properties of the built sm_100a step kernel (no GPU needed: cuobjdump reads the in-tree library).
Guards the design points DESIGN.md 5.1 relies on: TMA bulk copies + mbarrier for the time-table tiles, register
budget for 6 resident blocks per SM, FP64 seeds from MUFU, shared-memory K rows through immediate-offset LDS, the
instruction budget of the hot loop -- and that the FP64 instruction mix is the one the roofline's flop-per-step
calibration (profiles/flop_per_traj_step.json) was measured on."""
import json
import os
import re
import shutil
import subprocess

import pytest

from reentry_old_b200 import build as B

KERNEL = "_ZN2rb11k_propagateILi128ELi6ELb0EEEvNS_10StepParamsE"


@pytest.fixture(scope="module")
def sass():
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    lib = B.build()
    txt = subprocess.run(["cuobjdump", "-sass", "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    m = re.search(r"Function : " + KERNEL + r"\b(.*?)(?=\n\s*(?:Fatbin|Function : )|\Z)", txt, re.S)
    assert m, "step kernel not found in the library"
    return txt, m.group(1)


def _ops(body):
    ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", body)
    return {o: ops.count(o) for o in set(ops)}


def test_built_for_sm_100a_only(sass):
    txt, _ = sass
    archs = set(re.findall(r"arch = (sm_\w+)", txt))
    assert archs == {"sm_100a"}, archs


def test_step_kernel_uses_tma_bulk_copy_and_mbarrier(sass):
    _, body = sass
    assert "UBLKCP" in body                      # cp.async.bulk.shared::cluster.global (TMA bulk copy of a table tile)
    assert "SYNCS" in body                       # mbarrier arrive / try_wait
    assert "MUFU.RSQ64H" in body and "MUFU.RCP64H" in body   # FP64 seeds; no libdevice slow paths in the RHS


def test_step_kernel_register_and_instruction_budget(sass):
    txt, body = sass
    res = re.search(r"Function " + KERNEL + r":\s*\n\s*(.*)", txt)
    regs = int(re.search(r"REG:(\d+)", res.group(1)).group(1)) if res else None
    assert regs is not None and regs <= 80, regs          # 6 blocks x 128 threads x 80 registers fit the register file
    ops = _ops(body)
    fp64 = sum(ops.get(k, 0) for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    total = sum(ops.values())
    # seven inlined evaluations (the step's six + the launch's first): ~3.7 k instructions, 45 % of them FP64
    # (the flattened-loop kernel of round 1: 1.4 k instructions, 29 % FP64 -- and 10 % more executed per step)
    assert total <= 3900 and fp64 >= 0.44 * total, (fp64, total)
    # the stack holds at most two 4-byte words (loop counters, read and written once per STEP, not per stage); no state, K row
    # or 64-bit value lives there
    assert ops.get("LDL", 0) <= 4 and ops.get("STL", 0) <= 4
    assert not re.search(r"(LDL|STL)\.(64|128)", body)
    assert ops.get("CALL", 0) <= 12                            # only cold paths (64-bit division, NaN rows of a retiring lane) are calls


def test_flop_calibration_matches_the_built_kernel(sass):
    """bench.py's roofline multiplies trajectory-steps by profiles/flop_per_traj_step.json, an ncu measurement of ONE
    build of the kernel.  The file records that build's static FP64 instruction mix; if the kernel in the tree has a
    different mix the calibration is stale and this test says so (re-run tools/gpu_ncu_pass.sh, tools/calibrate_flop.py)."""
    _, body = sass
    ops = _ops(body)
    cal = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "flop_per_traj_step.json")))
    want = cal["static_fp64_of_calibrated_kernel"]
    have = {k: ops.get(k, 0) for k in ("DFMA", "DMUL", "DADD")}
    assert have == want, (have, want)


def test_step_kernel_shared_memory_accesses_have_immediate_offsets(sass):
    _, body = sass
    lds = re.findall(r"LDS(?:\.\d+)?\s+R\d+, \[(R\d+)(\+0x[0-9a-f]+)?\]", body)
    assert len(lds) >= 100                                # per evaluation: K rows, time-table entry, parked values, lookup tables
    with_imm = sum(1 for _, off in lds if off)
    assert with_imm >= 0.8 * len(lds)                     # one base register, offsets folded into the instruction
    assert body.count("LDG.E.64.CONSTANT") <= 4   # (only the copy of the lookup tables into shared memory at kernel start reads them from global memory)
