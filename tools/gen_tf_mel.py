#!/usr/bin/env python3
"""Straight-line mel projection for the thread-per-frame kernel (csrc/logmel_tf_kernel.cuh).

In that kernel one thread owns one frame: after stage 2 it holds the power of the 20 bins of a
row k1 (bins k1 + 20 j, folded to <= 200) in registers, and the n_mels running sums of ITS frame
in registers as well.  Which sums a bin feeds is a property of the filter bank's sparsity pattern
only, so for the two banks the reference uses (WhisperFeatureExtractor.mel_filters with 80 and
128 filters, transformers/models/whisper/feature_extraction_whisper.py:94-102) the projection is
emitted as straight-line FFMAs: no index arithmetic, no shared memory, no constant loads -- the
weights are immediate operands of the FFMAs (loading 394 weights through uniform registers cost
~450 extra instructions per 32 frames, ptxas runs out of uniform registers).  lm_create compares
lm_config.fbank with the baked table bit for bit; any other bank takes the table-driven kernels.

    python tools/gen_tf_mel.py > mlx8_ws_audio_transformer_b200/csrc/tf_mel_gen.cuh
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "lm_filters", os.path.join(ROOT, "mlx8_ws_audio_transformer_b200", "filters.py"))
filters = importlib.util.module_from_spec(spec)
spec.loader.exec_module(filters)

N, N1 = 400, 20


def row_bins(k1: int):
    """bin held by output j of stage-2 row k1 (codelets stage2_r20_half / stage2_c20 / stage2_c20_half)."""
    if k1 == 0:
        return [20 * j for j in range(11)]
    if k1 == 10:
        return [10 + 20 * j for j in range(10)]
    out = []
    for j in range(20):
        k = k1 + 20 * j
        out.append(k if k <= 200 else N - k)
    return out


EPILOGUE_COST = 5      # instructions per finished filter (log2, fma, store, 2 x min/max share) vs 1 per weight
STAGE2_BIAS = -170     # role A (row 0, row pairs 0-1) carries ~170 FEWER stage-2 instructions per tile than role B (pairs 2-4)


def split_point(fb, nm: int) -> int:
    """first filter of role B: both warps of a pair should get the same stage-2 + mel + epilogue work"""
    nnz = [(fb[:, m] != 0).sum() for m in range(nm)]
    cost = [int(n) + EPILOGUE_COST for n in nnz]
    best, best_d = 2, None
    for m0 in range(2, nm - 1, 2):      # even: the epilogue walks filters two at a time
        d = abs((sum(cost[:m0]) + STAGE2_BIAS) - sum(cost[m0:]))
        if best_d is None or d < best_d:
            best, best_d = m0, d
    return best


def emit_bank(nm: int) -> str:
    fb = filters.slaney_mel_filter_bank(201, nm).astype(np.float32)
    nz = [(k, m) for k in range(201) for m in range(nm) if fb[k, m] != 0.0]
    m0 = split_point(fb, nm)
    order = []
    lines = []
    for role in (0, 1):
        lo, hi = (0, m0) if role == 0 else (m0, nm)

        def fmas(src, k):
            out = []
            for m in range(lo, hi):
                if fb[k, m] != 0.0:
                    order.append((k, m))
                    out.append(f"  acc[{m - lo}] = __builtin_fmaf({src}, {float(fb[k, m]):.9e}f, acc[{m - lo}]);")
            return out

        # row 0: p[j] = power of bin 20 j
        body = []
        for j, k in enumerate(row_bins(0)):
            body += fmas(f"p[{j}]", k)
        lines.append(f"template <> LM_D void tf_mel_row0<{nm}, {role}>(const float (&p)[12], "
                     f"float (&acc)[TfMelPattern<{nm}>::MAXHALF]) {{   // {len(body)} weights")
        lines += body
        lines.append("}")
        # row pair q = rows (2q+1, 2q+2): pp[2j] = row 2q+1 output j, pp[2j+1] = row 2q+2 output j
        for q in range(5):
            body = []
            for half, k1 in enumerate((2 * q + 1, 2 * q + 2)):
                for j, k in enumerate(row_bins(k1)):
                    body += fmas(f"pp[{2 * j + half}]", k)
            lines.append(f"template <> LM_D void tf_mel_pair<{nm}, {q}, {role}>(const float (&pp)[40], "
                         f"float (&acc)[TfMelPattern<{nm}>::MAXHALF]) {{   // rows {2 * q + 1}, {2 * q + 2}: {len(body)} weights")
            lines += body
            lines.append("}")
    assert sorted(order) == sorted(nz), "every non-zero weight is used exactly once"
    bins = ", ".join(str(k) for k, _ in order)
    mels = ", ".join(str(m) for _, m in order)
    vals = ", ".join(f"0x{int(np.float32(fb[k, m]).view(np.uint32)):08x}u" for k, m in order)
    nnz_a = sum(1 for _, m in order if m < m0)
    head = [
        f"// ---- {nm} filters: {len(order)} non-zero weights; role A takes filters [0, {m0}) ({nnz_a} weights), "
        f"role B [{m0}, {nm}) ({len(order) - nnz_a} weights)",
        f"template <> struct TfMelPattern<{nm}> {{",
        f"  static constexpr int NNZ = {len(order)};",
        f"  static constexpr int M0 = {m0};                       // first filter of role B",
        f"  static constexpr int MAXHALF = {max(m0, nm - m0)};",
        f"  static constexpr unsigned short bin[NNZ] = {{{bins}}};",
        f"  static constexpr unsigned char mel[NNZ] = {{{mels}}};",
        f"  static constexpr unsigned bits[NNZ] = {{{vals}}};   // float32 images of the weights baked into the code",
        "};",
    ]
    return "\n".join(head + lines)


HEADER = '''// GENERATED by tools/gen_tf_mel.py -- do not edit by hand.
//
// Straight-line banded mel projection for logmel_tf_kernel.cuh.  Stage 2 leaves the power spectrum of
// a frame in tensor memory row by row: row 0 alone (p[j] = |X|^2 of bin 20 j), rows (2q+1, 2q+2) as
// packed pairs (pp[2j], pp[2j+1] = output j of the two rows; output j of row k1 is bin k1 + 20 j,
// folded to <= 200).  tf_mel_row0 / tf_mel_pair<NM, Q, R> add those contributions to the per-frame
// sums of the filters of role R (the two warps that share a frame tile split the filters at
// TfMelPattern<NM>::M0; acc is indexed from the role's first filter).  The weights are literals (float32 images in TfMelPattern<NM>::bits, checked against
// lm_config.fbank by lm_create).
#pragma once
#include "vec_ops.cuh"

namespace lm {

constexpr int kTfMaxNnz = 400;
template <int NM> struct TfMelPattern;
template <int NM, int R> LM_D void tf_mel_row0(const float (&p)[12], float (&acc)[TfMelPattern<NM>::MAXHALF]);
template <int NM, int Q, int R> LM_D void tf_mel_pair(const float (&pp)[40], float (&acc)[TfMelPattern<NM>::MAXHALF]);
'''


def main():
    print(HEADER)
    for nm in (80, 128):
        print(emit_bank(nm))
        print()
    print("}  // namespace lm")


if __name__ == "__main__":
    main()
