"""The mbarrier protocol of the persistent attention kernel (csrc/attention_pipe.cu), executed as a discrete-event model
under random interleavings (tools/sim_attn_pipe.py): no deadlock, no phase aliasing, no buffer hazard — and the model does
notice protocols that are deliberately broken.  Plus the address model of the halo-block conv form's row-shifted UMMA descriptors
(csrc/gemm.cu), the arithmetic behind the hardware fact recorded in profiles/r2_conv_halo.md."""
import importlib.util
import os
import random

import pytest

_spec = importlib.util.spec_from_file_location(
    "sim_attn_pipe", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "sim_attn_pipe.py"))
sim = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sim)


@pytest.mark.parametrize("prefix", [0, 1])
def test_protocol_is_clean_under_random_schedules(prefix):
    for seed in range(40):
        for njobs in (1, 2, 3, 5):
            sim.Sim(njobs, prefix, random.Random(1000 * seed + njobs)).run()


@pytest.mark.parametrize("broken", ["no_q_empty_wait", "no_kv_empty_wait", "no_s_empty_wait", "no_pv0_wait"])
def test_model_catches_broken_protocols(broken):
    caught = 0
    for seed in range(30):
        try:
            sim.Sim(4, 1, random.Random(seed), broken=broken).run()
        except AssertionError:
            caught += 1
    assert caught > 0


def test_halo_block_descriptor_addressing():
    """Address model of the halo-block conv form (csrc/gemm.cu, BRES 2 / 3; measured in profiles/r2_conv_halo.md): TMA lands
    the (18 x 16)-pixel block as 128-byte pixel records with SWIZZLE_128B, i.e. 16-byte chunk c of the record at byte address a
    is stored at chunk c ^ ((a >> 7) & 7); a K-major SW128 UMMA descriptor with start = block + dy*2048 + dx*128 + j*32 and
    SBO = 2048 reads row i of 8-row group g at start + g*SBO + i*128 and applies the same XOR to the ABSOLUTE address (base
    offset field 0).  Under that model every tap (dy, dx), K-slice j, tile row g and pixel i of the 16 x 8 tile reads exactly
    the channels [16 j, 16 j + 16) of halo pixel (g + dy, i + dx) — for ANY dx, although the 8-row groups then straddle two
    1024-byte swizzle atoms."""
    HALO_W, HALO_H, REC = 16, 18, 128
    base = 3 * 1024                                   # any 1024-byte-aligned shared-memory address
    smem = {}                                         # byte address of a 16-byte chunk -> (halo y, halo x, logical chunk)
    for hy in range(HALO_H):
        for hx in range(HALO_W):
            rec = base + (hy * HALO_W + hx) * REC
            for c in range(8):
                smem[rec + ((c ^ ((rec >> 7) & 7)) << 4)] = (hy, hx, c)
    assert len(smem) == HALO_H * HALO_W * 8           # the swizzle is a permutation inside each record

    def umma_read(start, sbo, g, i, kchunk):          # 16-byte chunk `kchunk` (0..1) of a 32-byte K-slice
        a = start + g * sbo + i * REC + kchunk * 16
        return (a & ~0x70) | ((((a >> 4) & 7) ^ ((a >> 7) & 7)) << 4)

    for dy in range(3):
        for dx in range(3):
            for j in range(4):
                start = base + dy * (HALO_W * REC) + dx * REC + j * 32
                for g in range(16):                   # tile row = 8-row group
                    for i in range(8):                # pixel of the tile row
                        for kc in range(2):
                            assert smem[umma_read(start, HALO_W * REC, g, i, kc)] == (g + dy, i + dx, 2 * j + kc)
