"""CPU-only check of the BUILT library's machine code (cuobjdump, no GPU): the hot kernels really are tcgen05 / TMEM /
TMA kernels, and the tensor-core issue loops stay free of the `ELECT / R2UR.BROADCAST / BRA.U.ANY` waterfall that a
per-thread role dispatch produces (it cost 107 clk per tcgen05.mma in the attention kernel)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sass():
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    from streamflow_b200.build import build_library
    so = build_library()
    text = subprocess.run([tool, "-sass", so], capture_output=True, text=True, check=True).stdout
    per_fn, fn = collections.defaultdict(list), None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
        elif fn and re.search(r"/\*[0-9a-f]{4}\*/", line):
            per_fn[fn].append(line)
    assert per_fn, "no SASS found: was the library built for sm_100a?"
    return per_fn


def kernel(sass, needle):
    hits = [v for k, v in sass.items() if needle in k]
    assert hits, f"kernel {needle} missing from libstreamcorr.so"
    return hits[0]


def count(lines, mnemonic):
    return sum(1 for l in lines if re.search(r"\b" + re.escape(mnemonic), l))


def test_attention_kernel_runs_on_cta_pairs(sass):
    k = kernel(sass, "16gma_stats_kernelE")
    assert count(k, "UTCHMMA.2CTA") >= 24       # cta_group::2 MMAs (the unrolled [hi | lo] schedule)
    assert count(k, "UTCBAR.2CTA.MULTICAST") >= 4 and count(k, "UTMALDG.3D.2CTA") >= 2 and count(k, "LDTM") >= 2
    # consecutive MMAs of one K step are issued back to back
    idx = [i for i, l in enumerate(k) if "UTCHMMA" in l]
    gaps = sorted(b - a for a, b in zip(idx, idx[1:]))
    assert gaps[len(gaps) // 2] <= 6, f"median distance between tcgen05.mma issues: {gaps[len(gaps) // 2]} instructions"


def test_aggregate_streams_with_bulk_copies_and_tcgen05(sass):
    k = kernel(sass, "20gma_aggregate_kernelIfE")
    assert count(k, "UTCHMMA") >= 4 and count(k, "UBLKCP") >= 3 and count(k, "UTMALDG") >= 1 and count(k, "LDTM") >= 2
    assert count(k, "BRA.U.ANY") == 0
    assert count(k, "RED.E") == 0 and count(k, "ATOMG") == 0         # no global reduction / atomic: split-K is gone


def test_lookup_is_a_warp_specialised_cp_async_gather(sass):
    """Loader warps (LDGSTS + mbarrier arrivals) and interpolation warps coupled by mbarriers: one block-wide barrier at
    start-up, none per work item."""
    k = kernel(sass, "21corr_lookup_ws_kernelILb0E")
    assert count(k, "LDGSTS.E.BYPASS.128") >= 16 and count(k, "UTCHMMA") == 0
    assert count(k, "SYNCS.PHASECHK") >= 2 and count(k, "BAR.SYNC") <= 1 and count(k, "STL") == 0
    assert not any("18corr_lookup_kernel" in name for name in sass)      # the two-barrier kernel is gone


def test_cross_check_lookup_stages_through_registers(sass):
    """The independently synchronised lookup kernel the parity tests compare the default one with bit for bit: LDG.128 into
    registers, STS.128 into ONE window buffer, plain block barriers, no cp.async at all, no spills."""
    k = kernel(sass, "22corr_lookup_reg_kernelILb0E")
    assert count(k, "LDG.E.NA.128") >= 12 and count(k, "STS.128") >= 9 and count(k, "LDGSTS") == 0
    assert count(k, "STL") == 0 and count(k, "LDL") == 0


def test_motion_encoder_entry_is_a_two_gemm_tcgen05_kernel(sass):
    """sf_pcblock_ffn1: both 1x1 convolutions on tcgen05 (4 + 4 MMA issue sites), weights by TMA, accumulators read back with
    tcgen05.ld, GELU pairs via FFMA2 / FMUL2 and one MUFU.RCP each, no issue waterfall, no spills."""
    k = kernel(sass, "19pcblock_ffn1_kernelIffE")
    assert count(k, "UTCHMMA") >= 8 and count(k, "UTMALDG") >= 3 and count(k, "LDTM") >= 3 and count(k, "MUFU.RCP") >= 16
    assert count(k, "FFMA2") >= 64 and count(k, "FMUL2") >= 48            # GELU pairs on the packed-fp32 pipe
    assert count(k, "BRA.U.ANY") == 0 and count(k, "STL") == 0 and count(k, "LDL") == 0


def test_correlation_gemm_default_runs_on_cta_pairs(sass):
    """Round 2: the default correlation GEMM is the cta_group::2 kernel (M = 256, B tile shared by the pair)."""
    k = kernel(sass, "21corr_gemm_pair_kernelE")
    assert count(k, "UTCHMMA.2CTA") >= 4 and count(k, "UTMALDG.3D.2CTA") >= 2 and count(k, "UTCBAR.2CTA.MULTICAST") >= 2
    assert count(k, "LDTM") >= 2 and count(k, "BRA.U.ANY") == 0
    assert count(k, "RED.E") == 0 and count(k, "ATOMG") == 0
    assert count(k, "STS") >= 8 and count(k, "ST.E.128 desc") == 0     # staging uses shared-space stores
    assert not any("16corr_gemm_kernelE" in name for name in sass)     # the single-CTA kernel of round 1 is gone


def test_aggregate_applies_wv_with_a_second_tensor_core_gemm(sass):
    """W_v is applied inside the streaming kernel: two groups of tcgen05.mma (key loop + epilogue GEMM on the staged
    hi / lo tiles), two TMA tensor loads (X tiles, W_v), and no per-iteration mma.sync projection kernel is left."""
    k = kernel(sass, "20gma_aggregate_kernelIfE")
    assert count(k, "UTCHMMA") >= 4 + 16          # 4 per key block + 16 of the second GEMM
    assert count(k, "UTMALDG") >= 3               # X ring + the two 64-column blocks of W_v
    assert count(k, "STS.U16") >= 32              # transposed fp16 hi / lo staging of the first GEMM's result
    assert not any("gma_proj_v_kernel" in name for name in sass)
    cast = kernel(sass, "15gma_cast_kernelIfE")
    assert count(cast, "HMMA") == 0 and count(cast, "F2FP") >= 4
