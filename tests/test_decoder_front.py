"""Decoder front half (reference decoder/src/newdecoder.cpp:212-290): sync-word correlation, frame
alignment, 180-degree phase fix, Viterbi r=1/2 k=7, NRZ-M.

CPU part: the oracle's convolutional code is pinned by the reference's own constants -- the four 64-bit sync words of
newdecoder.cpp:21-24 are the encoded attached sync marker 0x1ACFFC1D -- and its decoder inverts its encoder through noise.
GPU part (`-m gpu`): xrd_correlate / xrd_decoder_front_run against the oracle, bit for bit.
"""
import numpy as np
import pytest

ASM = 0x1ACFFC1D


def _bits(v, n):
    return np.array([(v >> (n - 1 - i)) & 1 for i in range(n)], np.uint8)


def _word(coded_bits):
    """the correlator's word: coded bit c is sent as a positive symbol for c = 0, which the correlator reads as bit 1"""
    v = 0
    for b in coded_bits:
        v = (v << 1) | (1 - int(b))
    return v


def make_stream(oracle, rng, lrit, n_frames, amp=48, sigma=20.0, invert_from=None, lead=777, gap_at=None):
    """n_frames CCSDS-style frames (sync marker + random bytes) -> [NRZ-M] -> r=1/2 k=7 -> BPSK soft bytes with noise"""
    payload = rng.integers(0, 256, (n_frames, 1020), dtype=np.uint8)
    frames = np.concatenate([np.tile(np.array([0x1A, 0xCF, 0xFC, 0x1D], np.uint8), (n_frames, 1)), payload], axis=1)
    bits = np.unpackbits(frames.reshape(-1))
    if not lrit:
        bits, _ = oracle.nrzm_encode(bits, 0)
    coded, _ = oracle.conv_encode(bits, 0)
    sym = np.where(coded == 0, amp, -amp).astype(np.float64)      # coded 0 = positive symbol
    if invert_from is not None:
        sym[invert_from * 16384:] *= -1                            # the Costas loop slipped by 180 degrees
    parts = [rng.normal(0, sigma, lead)]
    if gap_at is not None:
        cut = gap_at * 16384
        parts += [sym[:cut] + rng.normal(0, sigma, cut), rng.normal(0, sigma, 5000), sym[cut:] + rng.normal(0, sigma, len(sym) - cut)]
    else:
        parts += [sym + rng.normal(0, sigma, len(sym))]
    parts += [rng.normal(0, sigma, 3000)]
    soft = np.clip(np.rint(np.concatenate(parts)), -128, 127).astype(np.int8)
    return frames, soft


# ------------------------------------------------------------------ CPU: the oracle is pinned by the reference's constants
def test_sync_words_of_the_reference_are_the_encoded_sync_marker(oracle):
    asm = _bits(ASM, 32)
    coded, _ = oracle.conv_encode(asm, 0)
    assert _word(coded) == oracle.UW["lrit"][0]                    # LRIT_UW0
    assert _word(1 - coded) == oracle.UW["lrit"][1]                # LRIT_UW2: the same, 180 degrees away
    for start, uw in ((0, oracle.UW["hrit"][0]), (1, oracle.UW["hrit"][1])):   # HRIT: NRZ-M first, from 0 / from 1
        d, _ = oracle.nrzm_encode(asm, start)
        coded, _ = oracle.conv_encode(d, 0)
        assert _word(coded) == uw


def test_viterbi_inverts_the_encoder_through_noise(oracle):
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2, 8224, dtype=np.uint8)
    coded, _ = oracle.conv_encode(bits, 0)
    clean = np.where(coded == 0, 64, -64).astype(np.int8)
    out, ber = oracle.viterbi27(clean, len(bits))
    assert np.array_equal(np.unpackbits(out)[: len(bits)], bits) and ber == 0
    noisy = np.clip(np.rint(clean + rng.normal(0, 30, len(clean))), -128, 127).astype(np.int8)
    hard_errors = int(np.count_nonzero((noisy.view(np.uint8) >> 7) != coded))
    assert hard_errors > 100
    # read as signed symbols the decoder corrects all of them (the open end of the block is less protected) ...
    out, ber = oracle.viterbi27(noisy, len(bits), soft_mode=1)
    assert np.array_equal(np.unpackbits(out)[: len(bits) - 40], bits[:-40]) and abs(ber - hard_errors) < 40
    # ... and read raw, as the reference call chain hands them over (confidence mirrored within each half), it still
    # decodes at this noise level, with less margin
    mild = np.clip(np.rint(clean + rng.normal(0, 16, len(clean))), -128, 127).astype(np.int8)
    out, ber = oracle.viterbi27(mild, len(bits), soft_mode=0)
    assert np.array_equal(np.unpackbits(out)[: len(bits) - 40], bits[:-40])


@pytest.mark.parametrize("soft_mode", [0, 1])
def test_viterbi_is_maximum_likelihood_by_brute_force(oracle, soft_mode):
    """for short blocks every (start state, bit sequence) can be enumerated: the decoder's output must have the
    smallest total metric sum |u - 255 c| of all 64 x 2^n candidates (an independent statement of what the
    add-compare-select recursion computes; numpy, not the oracle's code)"""
    n = 10
    rng = np.random.default_rng(77 + soft_mode)
    parity = np.array([bin(i).count("1") & 1 for i in range(128)], dtype=np.int64)
    starts = np.arange(64, dtype=np.int64)[:, None]
    seqs = np.arange(1 << n, dtype=np.int64)[None, :]
    for trial in range(6):
        soft = rng.integers(-128, 128, 2 * n).astype(np.int8)
        if trial == 0:
            soft[:] = 0          # all ties
        u = ((127 - soft.astype(np.int64)) & 0xFF) if soft_mode else soft.view(np.uint8).astype(np.int64)
        sr = np.broadcast_to(starts, (64, 1 << n)).copy()
        total = np.zeros((64, 1 << n), np.int64)
        for t in range(n):
            bit = (seqs >> (n - 1 - t)) & 1
            sr = ((sr << 1) | bit) & 0x7F
            ca, cb = parity[sr & 0x4F], parity[sr & 0x6D]
            total += np.abs(u[2 * t] - 255 * ca) + np.abs(u[2 * t + 1] - 255 * cb)
        best = int(total.min())
        out, _ = oracle.viterbi27(soft, n, soft_mode=soft_mode)
        got = int("".join(str(int(b)) for b in np.unpackbits(out)[:n]), 2)
        assert int(total[:, got].min()) == best, "trial %d: decoded sequence has metric %d, the best is %d" % (
            trial, int(total[:, got].min()), best)


def test_correlator_finds_planted_words(oracle):
    rng = np.random.default_rng(2)
    data = np.clip(np.rint(rng.normal(0, 40, 16384)), -128, 127).astype(np.int8)
    for word_no, pos in ((0, 0), (1, 9000), (0, 16319)):
        d = data.copy()
        w = oracle.UW["lrit"][word_no]
        for k in range(64):
            d[pos + k] = 60 if (w >> (63 - k)) & 1 else -60       # word bit 1 <-> byte < 127 (a positive symbol)
        assert oracle.correlate(d, oracle.UW["lrit"]) == (64, pos, word_no)
    assert oracle.correlate(data[:64], oracle.UW["lrit"]) == (0, 0, 0)   # nothing to search


@pytest.mark.parametrize("lrit", [True, False])
@pytest.mark.parametrize("soft_mode", [0, 1])
def test_decoder_front_recovers_frames(oracle, lrit, soft_mode):
    rng = np.random.default_rng(3)
    frames, soft = make_stream(oracle, rng, lrit, 6, sigma=12.0 if soft_mode == 0 else 24.0, invert_from=3 if lrit else None,
                               gap_at=2)
    got, meta, consumed = oracle.DecoderFront(lrit, soft_mode).run(soft)
    # (the very first bits after start-up have no history to lean on: the encoder's start state is unknown to the decoder)
    assert len(got) == 6 and np.array_equal(got[1:], frames[1:]) and np.array_equal(got[0, 1:], frames[0, 1:])
    assert meta[0][0] == 777 and all(m[1] >= 46 for m in meta)
    if lrit:
        assert [m[2] for m in meta] == [0, 0, 0, 1, 1, 1]
    assert consumed <= len(soft) and len(soft) - consumed < 16384


# ------------------------------------------------------------------ GPU parity
@pytest.mark.gpu
def test_gpu_correlate_matches_the_oracle(gpu, xrd, oracle):
    rng = np.random.default_rng(4)
    for lrit in (True, False):
        f = xrd.DecoderFront(lrit)
        words = oracle.UW["lrit" if lrit else "hrit"]
        for n in (16384, 1024, 65, 64, 10, 100000):
            data = np.clip(np.rint(rng.normal(0, 40, n)), -128, 127).astype(np.int8)
            if n >= 1024:
                pos = int(rng.integers(0, n - 64))
                w = words[int(rng.integers(0, 2))]
                for k in range(64):
                    data[pos + k] = 50 if (w >> (63 - k)) & 1 else -50
                data[pos + 3] = -data[pos + 3]                     # one symbol in error
            assert f.correlate(data) == oracle.correlate(data, words), n
        data = np.full(4096, 127, np.int8)                         # exactly 127 reads as "0"; ties go to the first position
        assert f.correlate(data) == oracle.correlate(data, words)


@pytest.mark.gpu
@pytest.mark.parametrize("lrit", [True, False])
@pytest.mark.parametrize("soft_mode", [0, 1])
def test_gpu_decoder_front_matches_the_oracle(gpu, xrd, oracle, lrit, soft_mode):
    rng = np.random.default_rng(5 + lrit)
    frames, soft = make_stream(oracle, rng, lrit, 40, sigma=12.0 if soft_mode == 0 else 26.0, invert_from=17 if lrit else None,
                               gap_at=9, lead=12345)
    ref_frames, ref_meta, ref_cons = oracle.DecoderFront(lrit, soft_mode).run(soft)
    got, meta, cons = xrd.DecoderFront(lrit, soft_mode).run(soft)
    assert cons == ref_cons and len(got) == len(ref_frames) == 40
    np.testing.assert_array_equal(got, ref_frames)
    assert [(m.offset, m.correlation, m.word, m.bit_errors) for m in meta] == [tuple(int(v) for v in r) for r in ref_meta]
    # ... and the frames are the ones that were sent (the open end of a block is less protected: allow a few bits)
    wrong = np.unpackbits(got[:, 1:] ^ frames[:, 1:]).sum()
    assert wrong <= 8, wrong
    # the same stream in two calls: the unconsumed tail is fed again, the previous frame's soft tail is carried
    a_ref, b_ref = oracle.DecoderFront(lrit, soft_mode), xrd.DecoderFront(lrit, soft_mode)
    cut = 300000
    f1, m1, c1 = a_ref.run(soft[:cut])
    f2, m2, c2 = a_ref.run(soft[c1:])
    g1, n1, d1 = b_ref.run(soft[:cut])
    g2, n2, d2 = b_ref.run(soft[d1:])
    assert (c1, c2) == (d1, d2)
    np.testing.assert_array_equal(np.concatenate([g1, g2]), np.concatenate([f1, f2]))
    assert np.unpackbits(np.concatenate([g1, g2])[:, 1:] ^ frames[:, 1:]).sum() <= 8


@pytest.mark.gpu
def test_gpu_decoder_front_on_the_demodulators_own_bytes(gpu, xrd, oracle):
    """noise-only soft bytes straight from xrd_demod_batch_i8 (no frames inside): nothing passes the correlation
    threshold, like the oracle; and a large random stream walks identically"""
    rng = np.random.default_rng(9)
    soft = np.clip(np.rint(rng.normal(0, 50, 3_000_000)), -128, 127).astype(np.int8)
    ref = oracle.DecoderFront(True).run(soft)
    got = xrd.DecoderFront(True).run(soft)
    assert got[2] == ref[2] and len(got[0]) == len(ref[0])
    np.testing.assert_array_equal(got[0], ref[0])
