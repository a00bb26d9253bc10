"""Static guards on the BUILT library (no GPU needed: cuobjdump reads the sm_100a code here).

Two properties the measured performance depends on and that a recompilation can silently lose:
  * the hot instantiations fit their register budget without spilling (80 registers = 3 CTAs/SM for the edge kernels,
    <= 128 for the two-CTA GEMM);
  * in the gather loops of the edge kernels every load of a group of edges is ISSUED before the first FMA that consumes one
    (memory-level parallelism).  ptxas once interleaved loads and FMAs under register pressure and the backward pass lost
    12 %; this test would have caught it without a GPU.
Also checks the SASS evidence the design claims: tcgen05 MMAs, TMA loads/stores and FP64 tensor-core MMAs in the GEMM,
bulk copies in the backward.
"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gat-pytorch_b200", "libgat_b200.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.isfile(LIB), reason="needs cuobjdump and the built library")

FWD = "edge_fwd_kernelILi32ELi2ELi4ELb0ELb0E"                      # products hidden layer, short rows
BWD = "edge_bwd_main_kernelILi32ELi2ELi4ELb0ELb1ELb1ELb0ELb0E"     # fused backward, FULL rows
BWD_GS = "edge_bwd_main_kernelILi32ELi2ELi4ELb0ELb1ELb0ELb1ELb0E"  # head-mean layer, staged rows
BWD_HM = "edge_bwd_hm4_kernelILi3ELb1EE"                             # head-mean layer (4 heads), short rows: lane = (edge slot, head, quarter)
ROWDOT = "edge_bwd_rowdot_kernelILi4ELb0EE"                         # per-node S = <dOut, out> pass, plain instantiation (3 CTAs/SM)
GEMM_NT = "gemm_tc_kernelILi128ELb0E"
GEMM_PAIR = "gemm_pair_kernel"


@pytest.fixture(scope="module")
def resources():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    res, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and name:
            res[name] = (int(m.group(1)), int(m.group(2)))
    return res


def _sass(function_substring):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", _mangled(function_substring), LIB], capture_output=True, text=True).stdout
    return [m.group(1).strip() for m in re.finditer(r"^\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", out, re.M)]


_NAMES = None


def _mangled(sub):
    global _NAMES
    if _NAMES is None:
        out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
        _NAMES = re.findall(r"Function (\S+):", out)
    hits = [n for n in _NAMES if sub in n]
    assert len(hits) == 1, (sub, hits)
    return hits[0]


def _op(ins):
    t = ins.split()
    return (t[1] if t[0].startswith("@") else t[0])


@pytest.mark.parametrize("sub,max_regs", [(FWD, 80), (BWD, 80), (BWD_GS, 80), (BWD_HM, 80), (ROWDOT, 80), (GEMM_NT, 128), (GEMM_PAIR, 200)])
def test_hot_kernels_fit_their_register_budget_without_spills(resources, sub, max_regs):
    regs, stack = resources[_mangled(sub)]
    # the pair GEMM keeps one loop counter on the stack (8 bytes, outside the hot loops); everything else must be spill-free
    assert regs <= max_regs and stack <= (16 if sub == GEMM_PAIR else 0), (sub, regs, stack)


@pytest.mark.parametrize("sub", [FWD, BWD])
def test_gather_loads_of_a_group_are_issued_before_the_first_fma(sub):
    ops = [_op(i) for i in _sass(sub)]
    # the full-group block: a run containing 8 x LDG.E.128 (4 edges x 2 chunks) with nothing but address arithmetic between
    # them, followed by the FMAs
    best = 0
    i = 0
    while i < len(ops):
        if ops[i].startswith("LDG.E.128"):
            j, n_ldg = i, 0
            while j < len(ops) and not ops[j].startswith(("FFMA", "FMUL", "BRA", "STS")):
                n_ldg += ops[j].startswith("LDG.E.128")
                j += 1
            best = max(best, n_ldg)
            i = j
        else:
            i += 1
    assert best >= 8, f"{sub}: at most {best} of the 8 gather loads of a group are in flight before the first FMA"


def test_sass_carries_the_instructions_the_design_claims():
    gemm = " ".join(_op(i) for i in _sass(GEMM_NT))
    assert "UTCHMMA" in gemm or "UTCMMA" in gemm            # tcgen05.mma
    assert "UTMALDG" in gemm and "UTMASTG" in gemm          # TMA tile loads, TMA stores of the output tile
    assert "DMMA" in gemm                                   # fused score epilogue on the FP64 tensor cores
    assert "LDTM" in gemm                                   # tcgen05.ld (TMEM -> registers)
    pair = [_op(i) for i in _sass(GEMM_PAIR)]
    assert any(o.startswith("UTCHMMA.2CTA") for o in pair)              # tcgen05.mma.cta_group::2
    assert any(o.startswith("UTCBAR.2CTA.MULTICAST") for o in pair)     # commit to both CTAs of the pair
    assert any(o.startswith("STTM") for o in pair) and any(o.startswith("LDTM") for o in pair)   # A operand written to / accumulators read from TMEM
    assert any(o.startswith("UTMALDG") for o in pair) and any(o.startswith("UTMASTG") for o in pair) and any(o.startswith("UTMAPF.L2") for o in pair)
    # the issue loop is warp-uniform: no ELECT / R2UR.BROADCAST re-derivation of the uniform operands between the MMAs
    first = next(i for i, o in enumerate(pair) if o.startswith("UTCHMMA"))
    last = max(i for i, o in enumerate(pair) if o.startswith("UTCHMMA"))
    assert not any(o.startswith("R2UR.BROADCAST") for o in pair[first:last]), "MMA issue loop lost its warp-uniform form"
    assert last - first <= 12 * 8, "more than ~8 instructions per MMA in the issue loop"
    bwd = " ".join(_op(i) for i in _sass(BWD))
    assert "UBLKCP" in bwd or "BLKCP" in bwd                # cp.async.bulk row push (partitioned runs)
    gs = " ".join(_op(i) for i in _sass(BWD_GS))
    assert "LDGSTS" in gs                                   # cp.async staging of the narrow shared rows
    hm = [_op(i) for i in _sass(BWD_HM)]
    assert any(o.startswith("LDGSTS") for o in hm)          # the four-head head-mean kernel stages its gathers the same way
    assert not any(o.startswith("BAR") for o in hm)         # and is warp-autonomous: no CTA barrier anywhere
