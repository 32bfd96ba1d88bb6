"""GPU parity tests of the hashing path through the C ABI.  Bar: crops, the 16x16x16 resized cube and the hash
words are BIT-EXACT against the CPU oracle (the DCT runs in f64 with the oracle's operation order, so there is no
epsilon band); the oracle itself is anchored as tests/test_oracle_hashing.py describes."""
import json
import os

import numpy as np
import pytest
import torch

import vid_dup_finder_lib_b200 as vdf
from oracle import vdf_oracle as o
from tests import synth
from tests.test_oracle_letterbox import KATS
from vid_dup_finder_lib_b200 import _ffi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ctx():
    return _ffi.default_context()


def gpu_hash(ctx, stacks: np.ndarray, cropdetect=1):
    """stacks [n,16,h,w] u8 (host) -> hash, status, crop via vdf_hash_stacks"""
    n, t, h, w = stacks.shape
    return ctx.hash_stacks(stacks.reshape(-1), _ffi.make_descs(n, w, h, t), cropdetect)


def gpu_small(ctx, stacks: np.ndarray, cropdetect=1):
    n, t, h, w = stacks.shape
    d = torch.from_numpy(stacks).cuda()
    out = torch.empty((n, 16, 16, 16), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    crop = ctx.hash_stacks_small_device(d.data_ptr(), _ffi.make_descs(n, w, h, t), cropdetect, out.data_ptr())
    return out.cpu().numpy(), crop


def oracle_all(stacks: np.ndarray, cropdetect=1):
    H, S, C, M = [], [], [], []
    for s in stacks:
        st, h, crop, small = o.hash_stack(s, cropdetect)
        H.append(h), S.append(st), C.append(crop), M.append(small)
    return np.stack(H), np.array(S), np.array(C), np.stack(M)


def test_letterbox_kat_images(ctx):
    """the reference's KAT images (video_frames_gray.rs:225-458) at the hot path's fixed AnyColour(16) setting,
    plus the oracle's verdict on the same images with frame 8 differing from frame 0"""
    for name, image, _, _, _ in KATS:
        h, w = image.shape
        big = np.kron(image, np.ones((7, 9), np.uint8))  # keep the geometry, give the resize something to chew
        for im in (image, big):
            st = np.broadcast_to(im, (1, 16) + im.shape).copy()
            _, status, crop = gpu_hash(ctx, st, 1)
            want = o.hash_stack(st[0], 1)
            assert status[0] == want[0] == 0 and tuple(crop[0]) == want[2], name
            assert tuple(crop[0]) == o.letterbox_frame(im, o.LB_ANYCOLOUR, 16), name


@pytest.mark.parametrize("variant", [-1, 0, 1, 2], ids=["fused", "imma4", "general", "imma8"])
@pytest.mark.parametrize("w,h", [(64, 48), (256, 144), (321, 203), (640, 360), (854, 480), (1280, 720), (1920, 1080)])
def test_crop_cube_and_hash_are_bit_exact(ctx, w, h, variant):
    """-1 (the default): hash_fused_kernel, one persistent launch per call; 0 / 2: the per-frame tensor-core (IMMA) resize kernels
    behind the letterbox scan kernels where rows are 16-byte aligned, general kernel elsewhere; 1: general kernel only"""
    ctx.set_option("hash_fused", 1 if variant < 0 else 0)
    ctx.set_option("hash_variant", max(variant, 0))
    try:
        _check_bit_exact(ctx, w, h)
    finally:
        ctx.set_option("hash_variant", 0)
        ctx.set_option("hash_fused", 1)


def _check_bit_exact(ctx, w, h):
    n = 6 if w * h > 500_000 else 12
    stacks = synth.frame_stacks(n, w, h, seed=w * 7 + h).numpy()
    stacks[1, :, : h // 9, :] = 16            # letterbox
    stacks[1, :, h - h // 11:, :] = 20
    stacks[2, :, :, : w // 10] = 235          # pillarbox, bright
    stacks[2, :, :, w - w // 13:] = 235
    stacks[3, 8] = stacks[3, 0]
    stacks[3, 8, : h // 6, :] = 0             # only frame 8 has a bar -> per-side min removes it
    stacks[4] = stacks[4, 0]                  # static: every t>0 coefficient is an exact zero
    stacks[5] = 77                            # uniform: every strip is "letterbox" -> no crop at all
    want_h, want_s, want_c, want_m = oracle_all(stacks)
    got_m, got_c2 = gpu_small(ctx, stacks)
    got_h, got_s, got_c = gpu_hash(ctx, stacks)
    assert np.array_equal(got_s, want_s)
    assert np.array_equal(got_c, want_c) and np.array_equal(got_c2, want_c)
    assert np.array_equal(got_m, want_m), "resized cube differs"
    assert np.array_equal(got_h, want_h), "hash words differ"
    # (smooth content can itself pass the strip test at small sizes, so only the planted geometry is asserted)
    assert want_c[1][2] >= h // 9 and want_c[2][0] >= w // 10 and want_c[3][2] < h // 6 and tuple(want_c[5]) == (0, 0, 0, 0)
    bits = np.unpackbits(got_h[4].view(np.uint8), bitorder="little")
    assert bits[100:].sum() == 0  # static stack


def test_letterbox_noisy_and_wide_bars(ctx):
    """the scan's shortcuts against the oracle: bars whose strips span more than the tolerance but are letterbox by the 90 %
    rule (the one-strip probe must fall through to the full histograms), bars wider than the 8 speculative panels (the serial
    tail), bars that differ between frame 0 and frame 8, a picture that starts mid-panel, sizes that are not multiples of 32"""
    rng = np.random.default_rng(5)
    cases = []
    for (w, h) in [(640, 360), (1000, 562), (1920, 1080)]:
        base = synth.frame_stacks(6, w, h, seed=w + h, bars=False).numpy()
        a = base[0]  # noisy letterbox: 6 % of the bar pixels are outliers (range 200, still > 90 % within 16 of the mode)
        t, b = h // 7 + 3, h // 9 + 1
        bar = np.where(rng.random((16, t, w)) < 0.06, 216, 16 + rng.integers(-3, 4, (16, t, w))).astype(np.uint8)
        a[:, :t, :] = bar
        a[:, h - b:, :] = np.where(rng.random((16, b, w)) < 0.06, 216, 18).astype(np.uint8)
        c = base[1]  # noisy pillarbox
        l, r = w // 8 + 5, w // 11
        c[:, :, :l] = np.where(rng.random((16, h, l)) < 0.05, 180, 12 + rng.integers(-2, 3, (16, h, l))).astype(np.uint8)
        c[:, :, w - r:] = np.where(rng.random((16, h, r)) < 0.05, 0, 235).astype(np.uint8)
        d = base[2]  # bars beyond 8 panels of 32 on every side that can hold them
        d[:, : min(300, h // 3), :] = 16
        d[:, :, : min(290, w // 3)] = 16
        d[:, :, w - min(270, w // 4):] = 20
        e = base[3]  # frame 8 has a narrower bar than frame 0: the per-side minimum
        e[:, :70, :] = 0
        e[8, 40:70, :] = base[4][8, 40:70, :]
        f = base[5]  # clean bar, one bright logo pixel row inside it
        f[:, :50, :] = 16
        f[:, 20, : w // 30] = 250
        cases += [a, c, d, e, f]
    for k, st in enumerate(cases):
        want = o.hash_stack(st, 1)
        got_h, got_s, got_c = gpu_hash(ctx, st[None], 1)
        assert got_s[0] == want[0] == 0 and tuple(int(v) for v in got_c[0]) == tuple(want[2]), (k, got_c[0], want[2])
        assert np.array_equal(got_h[0], want[1]), k
    crops = [tuple(o.hash_stack(st, 1)[2]) for st in cases[:5]]
    assert crops[0][2] >= 360 // 7 and crops[1][0] >= 640 // 8 and crops[2][0] > 200 and crops[3][2] == 40, crops


def test_fused_kernel_machinery(ctx):
    """hash_fused_kernel's own moving parts against the oracle: (a) 700 stacks of one-tile frames -- the producers run four frames ahead
    of the consumers, the slot ring, the put-aside list, the double-buffered vertical sums and the finalizer all turn over every
    few hundred cycles, several frames per block are in flight at once; (b) one call mixing stacks in error, stacks whose rows are
    not 16-byte aligned (general kernel behind the fused launch), bars, and more stacks than blocks; (c) frames beyond the
    one-round-trip strip-0 path (wider than 2048, taller than 1152) and beyond the per-frame kernels' shared-memory budget"""
    rng = np.random.default_rng(77)
    # (a)
    n, w, h = 700, 96, 64
    st = synth.frame_stacks(n, w, h, seed=99).numpy()
    for s in range(0, n, 5):
        st[s, :, : int(rng.integers(1, 20)), :] = 16
        st[s, :, :, w - int(rng.integers(1, 30)):] = 20
    want_h, want_s, want_c, _ = oracle_all(st)
    got_h, got_s, got_c = gpu_hash(ctx, st)
    assert np.array_equal(got_c, want_c) and np.array_equal(got_s, want_s) and np.array_equal(got_h, want_h)
    # (b)
    sizes = [(320, 180), (333, 187), (640, 360), (48, 48), (320, 180), (1000, 562), (320, 180)] * 30
    descs = np.zeros(len(sizes), dtype=_ffi.STACK_DESC_DTYPE)
    bufs, stacks, off = [], [], 0
    for k, (w, h) in enumerate(sizes):
        one = synth.frame_stacks(1, w, h, seed=1000 + k)[0].numpy()
        if k % 4 == 1:
            one[:, : h // 8, :] = 16
            one[:, h - h // 10:, :] = 18
        if k % 7 == 3:
            one[:, :, : w // 9] = 12
        pitch = w + (0 if k % 3 else 7)  # every third stack: rows not 16-byte aligned
        n_fr = 16 if k % 11 else 9       # every eleventh: NotEnoughFrames
        buf = np.zeros((n_fr, h, pitch), np.uint8)
        buf[:, :, :w] = one[:n_fr]
        descs[k] = (off, pitch * h, w, h, pitch, n_fr, _ffi.STACK_FLAG_MIXED_SIZES if k % 13 == 5 else 0, 0)
        bufs.append((off, buf.reshape(-1)))
        stacks.append(one)
        off += (buf.size + 15) // 16 * 16
    host = np.zeros(off, np.uint8)
    for o0, b1 in bufs:
        host[o0:o0 + b1.size] = b1
    got_h, got_s, got_c = ctx.hash_stacks(host, descs, 1)
    for k, one in enumerate(stacks):
        if k % 13 == 5:
            assert got_s[k] == _ffi.STACK_VIDPROC and not got_h[k].any(), k
        elif k % 11 == 0:
            assert got_s[k] == _ffi.STACK_NOT_ENOUGH_FRAMES and not got_h[k].any(), k
        else:
            w_st, w_h, w_crop, _ = o.hash_stack(one, 1)
            assert got_s[k] == w_st == 0 and tuple(got_c[k]) == w_crop and np.array_equal(got_h[k], w_h), (k, sizes[k])
    # (c)
    for (w, h) in [(2304, 400), (640, 1300), (2560, 1440), (256, 9000)]:
        st = synth.frame_stacks(3, w, h, seed=w + h).numpy()
        st[1, :, : h // 7, :] = 16
        st[1, :, h - h // 9:, :] = 16
        st[2, :, :, : w // 6] = 235
        want_h, want_s, want_c, _ = oracle_all(st)
        got_h, got_s, got_c = gpu_hash(ctx, st)
        assert np.array_equal(got_c, want_c) and np.array_equal(got_s, want_s) and np.array_equal(got_h, want_h), (w, h)


def test_chunked_overlapped_pipeline_matches_serial(ctx):
    """the fused kernel (cold tables: sizes met for the first time take a second pass; warm; DCT in the kernel or on its own) and the
    per-frame pipeline it replaced: resize jobs built on the device from the crops (sizes met for the first time take a second pass),
    the letterbox scan of chunk k+1 on a second stream beside the resize of chunk k, DCT + pack fused into the resize kernel.
    Every combination of the knobs gives the oracle's hashes, on a fresh context (cold tables) and again (warm)."""
    n, w, h = 160, 256, 144
    stacks = synth.frame_stacks(n, w, h, seed=4321).numpy()
    rng = np.random.default_rng(5)
    for s in range(0, n, 3):  # many distinct crop sizes: bars of different widths on different sides
        t, b, l = int(rng.integers(0, 30)), int(rng.integers(0, 30)), int(rng.integers(0, 40))
        stacks[s, :, :t, :] = 16
        stacks[s, :, h - b:, :] = 18
        stacks[s, :, :, :l] = 14
    want_h, want_s, want_c, _ = oracle_all(stacks)
    fresh = _ffi.Context(ctx.device)
    try:
        for fused, overlap, chunks, fuse in ((1, 0, 1, 1), (1, 0, 1, 1), (1, 0, 1, 0), (0, 1, 4, 1), (0, 1, 4, 1), (0, 0, 1, 1), (0, 0, 4, 0),
                                             (0, 1, 2, 0), (0, 0, 1, 0)):
            fresh.set_option("hash_fused", fused)
            fresh.set_option("hash_overlap", overlap)
            fresh.set_option("hash_chunks", chunks)
            fresh.set_option("hash_fuse_dct", fuse)
            got_h, got_s, got_c = gpu_hash(fresh, stacks)
            assert np.array_equal(got_c, want_c), (fused, overlap, chunks, fuse)
            assert np.array_equal(got_s, want_s) and np.array_equal(got_h, want_h), (fused, overlap, chunks, fuse)
    finally:
        fresh.close()


def test_cropdetect_none(ctx):
    stacks = synth.frame_stacks(4, 320, 180, seed=3).numpy()
    stacks[:, :, :20, :] = 16
    want_h, _, want_c, _ = oracle_all(stacks, 0)
    got_h, got_s, got_c = gpu_hash(ctx, stacks, 0)
    assert np.array_equal(got_h, want_h) and not got_c.any() and not got_s.any()
    assert not np.array_equal(got_h, gpu_hash(ctx, stacks, 1)[0])


def test_dct_threshold_pack_bit_exact(ctx):
    rng = np.random.default_rng(11)
    cubes = rng.integers(0, 256, (300, 16, 16, 16), dtype=np.uint8)
    cubes[0] = 128                                   # all-zero input: every coefficient 0.0 -> empty hash
    cubes[1] = cubes[1, 0]                           # static
    cubes[2] = np.concatenate([cubes[2, :, :, :8], cubes[2, :, :, 7::-1]], axis=2)  # mirror-symmetric in x
    cubes[3] = 255
    cubes[4] = rng.integers(126, 131, (16, 16, 16))  # tiny amplitudes
    got = ctx.hash_from_small(cubes)
    want = np.stack([o.hash_from_small(c) for c in cubes])
    assert np.array_equal(got, want)
    assert not got[0].any() and (got[:, 15] >> np.uint64(40)).max() == 0  # pad bits 1000..1023 stay zero


def test_reference_example_clips(ctx):
    """examples/example.rs:77-82 on the decodable clips, end to end on the GPU: golden words, then search()."""
    z = np.load(os.path.join(GOLD, "ref_clips_gray.npz"))
    gold = json.load(open(os.path.join(GOLD, "ref_clips_hashes.json")))
    keys = [n.replace(".", "_") for n in z["names"]] + ["cat_1_letterboxed"]
    b = vdf.VideoHashBuilder()
    durs = list(z["durations"]) + [45]
    res = b.hash_many([list(z[k]) for k in keys], keys, durs)
    for k, r in zip(keys, res):
        assert isinstance(r, vdf.VideoHash)
        assert [format(w, "016x") for w in r.hash] == gold[k]["hash_words_hex"], k
    groups = vdf.search(res, vdf.DEFAULT_SEARCH_TOLERANCE)
    assert sorted(sorted(g.duplicates()) for g in groups) == [["cat_1_letterboxed", "cat_1_mp4", "cat_3_webm"],
                                                              ["dog_1_mp4", "dog_3_webm"]]
    # lib.rs:26-62 doc-test: cat.1 + cat.3 match, dog.3 does not
    sub = [res[0], res[1], res[3]]
    g2 = vdf.search(sub, vdf.DEFAULT_SEARCH_TOLERANCE)
    assert len(g2) == 1 and g2[0].len() == 2


def test_errors_mirror_the_reference(ctx):
    fr = [synth.smooth_frame(96, 64, s) for s in range(16)]
    b = vdf.VideoHashBuilder()
    ok = b.hash_frames(fr, "a.mp4", 12)
    assert ok.src_path == "a.mp4" and ok.duration == 12
    assert np.array_equal(ok.words(), o.hash_stack(np.stack(fr), 1)[1])
    with pytest.raises(vdf.NotEnoughFrames):  # dct_3d.rs:47-52
        b.hash_frames(fr[:15], "a", 1)
    with pytest.raises(vdf.NotEnoughFrames):
        b.hash_frames([], "a", 1)
    bad = list(fr)
    bad[5] = synth.smooth_frame(96, 66, 5)
    with pytest.raises(vdf.VidProc):  # video_hash_builder.rs:169-186
        b.hash_frames(bad, "a", 1)
    more = b.hash_frames(fr + fr[:4], "a.mp4", 12)  # take(16), video_hash_builder.rs:164
    assert more.hash == ok.hash
    mixed = b.hash_many([fr, fr[:3], bad, fr], ["0", "1", "2", "3"], [1, 2, 3, 4])
    assert isinstance(mixed[0], vdf.VideoHash) and isinstance(mixed[1], vdf.NotEnoughFrames)
    assert isinstance(mixed[2], vdf.VidProc) and mixed[3].hash == ok.hash and mixed[3].duration == 4
    no_crop = vdf.VideoHashBuilder(vdf.CreationOptions(cropdetect=vdf.Cropdetect.NONE)).hash_frames(fr, "a", 1)
    assert np.array_equal(no_crop.words(), o.hash_stack(np.stack(fr), 0)[1])


def test_ragged_batch_and_pitched_frames(ctx):
    """stacks of different sizes in one call, frames with row padding (pitch > width), pinned and pageable hosts"""
    sizes = [(160, 90), (33, 17), (640, 360), (16, 16), (7, 5), (200, 200)]
    stacks = [synth.frame_stacks(1, w, h, seed=w)[0].numpy() for w, h in sizes]
    descs = np.zeros(len(sizes), dtype=_ffi.STACK_DESC_DTYPE)
    chunks, off = [], 0
    for k, (st, (w, h)) in enumerate(zip(stacks, sizes)):
        pitch = w + (k % 3) * 5
        buf = np.zeros((16, h, pitch), np.uint8)
        buf[:, :, :w] = st
        descs[k] = (off, pitch * h, w, h, pitch, 16, 0, 0)
        chunks.append(buf.reshape(-1))
        off += buf.size
    host = np.concatenate(chunks)
    want = np.stack([o.hash_stack(st, 1)[1] for st in stacks])
    got, status, _ = ctx.hash_stacks(host, descs, 1)
    assert not status.any() and np.array_equal(got, want)
    pinned = torch.from_numpy(host).pin_memory()
    got2, _, _ = ctx.hash_stacks(pinned.numpy(), descs, 1)
    assert np.array_equal(got2, want)


def test_hash_pipeline_many_producers(ctx):
    """decode threads push stacks, the library batches them (here: tiny batches, so producers block and
    batches alternate), results come back tagged; every hash / error equals the direct call's and the oracle's."""
    import threading

    from vid_dup_finder_lib_b200.pipeline import HashPipeline

    sizes = [(160, 90), (64, 48), (320, 180), (33, 17)]
    stacks = {}
    for tag in range(48):
        w, h = sizes[tag % len(sizes)]
        st = synth.frame_stacks(1, w, h, seed=100 + tag)[0].numpy()
        if tag % 11 == 5:
            stacks[tag] = list(st[:9])  # NotEnoughFrames
        elif tag % 11 == 7:
            stacks[tag] = list(st[:8]) + list(synth.frame_stacks(1, w + 2, h, seed=tag)[0].numpy()[:8])  # VidProc
        elif tag % 11 == 9:
            stacks[tag] = [np.asfortranarray(f) for f in st] + [st[0]] * 3  # > 16 frames (take(16)), non-contiguous arrays
        else:
            stacks[tag] = list(st)
    got = {}
    with HashPipeline(vdf.CreationOptions(), ctx=ctx, max_batch_stacks=5, batch_bytes=3 * 16 * 320 * 180 + 4096) as p:
        def producer(k):
            for tag in range(k, 48, 4):
                p.push(tag, stacks[tag])

        def collector():
            while len(got) < 48:
                for tag, val, crop in p.results(wait=True):
                    got[tag] = (val, crop)
                if not th_alive():
                    p.flush()

        threads = [threading.Thread(target=producer, args=(k,)) for k in range(4)]
        th_alive = lambda: any(t.is_alive() for t in threads)
        for t in threads:
            t.start()
        c = threading.Thread(target=collector)
        c.start()
        for t in threads:
            t.join(timeout=120)
        p.flush()
        c.join(timeout=120)
        assert not c.is_alive() and len(got) == 48
    for tag, (val, crop) in got.items():
        fr = stacks[tag]
        if tag % 11 == 5:
            assert isinstance(val, vdf.NotEnoughFrames)
        elif tag % 11 == 7:
            assert isinstance(val, vdf.VidProc)
        else:
            cube = np.stack(fr[:16])
            _, want_hash, want_crop, _ = o.hash_stack(cube, 1)
            assert np.array_equal(val, want_hash), tag
            assert tuple(int(x) for x in crop) == tuple(int(x) for x in want_crop), tag


def test_full_size_1080p_batch(ctx):
    """BASELINE config shape: 1080p stacks resident in HBM, generated on the device; parity of a sample against
    the oracle (which needs ~0.3 s per stack), invariants on all of them."""
    n = 48
    d = synth.frame_stacks(n, 1920, 1080, seed=synth.SEED, device="cuda")
    out = torch.zeros((n, 16), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    st, crop = ctx.hash_stacks_device(d.data_ptr(), _ffi.make_descs(n, 1920, 1080), 1, out.data_ptr())
    got = out.cpu().numpy().view(np.uint64)
    assert not st.any()
    assert (got[:, 15] >> np.uint64(40)).max() == 0
    assert (crop[:, 0] + crop[:, 1] < 1920).all() and (crop[:, 2] + crop[:, 3] < 1080).all() and crop.any()
    for s in range(0, n, 6):
        w_st, w_h, w_crop, _ = o.hash_stack(d[s].cpu().numpy(), 1)
        assert w_st == 0 and tuple(crop[s]) == w_crop and np.array_equal(got[s], w_h), s
    # idempotence: hashing the same resident stacks again gives the same words
    out2 = torch.zeros_like(out)
    ctx.hash_stacks_device(d.data_ptr(), _ffi.make_descs(n, 1920, 1080), 1, out2.data_ptr())
    assert torch.equal(out, out2)


def _video_in_a_frame(w, h, seed, box=None, second=None, stretch=False):
    """a bright static surround with a darker picture that moves inside `box` = (x, y, bw, bh) -- what Cropdetect::Motion is for"""
    rng = np.random.default_rng(seed)
    st = np.full((16, h, w), 235, np.uint8)
    st += rng.integers(0, 3, (1, h, w), dtype=np.uint8)  # static texture: no motion outside the boxes
    for bx in (box, second):
        if bx is None:
            continue
        x, y, bw, bh = bx
        yy, xx = np.mgrid[0:bh, 0:bw]
        for t in range(16):
            pic = 90 + 60 * np.sin(2 * np.pi * (xx / bw * 2 + t / 16.0)) * np.cos(2 * np.pi * (yy / bh + t / 8.0)) + rng.integers(-6, 7, (bh, bw))
            st[t, y:y + bh, x:x + bw] = np.clip(pic, 1, 200).astype(np.uint8)
    if stretch:  # neither 0 nor 255 anywhere: the contrast stretch applies
        st = (20 + st.astype(np.uint16) * 180 // 255).astype(np.uint8)
    return st


def test_motion_crop_matches_oracle(ctx):
    """Cropdetect::Motion on the GPU (csrc/motion.cu) against oracle/motioncrop_oracle.py: crop and hash words, on stacks with
    one moving picture, two (the second pass and the selection rule), none (letterbox fallback), with and without the contrast
    stretch, below and above the 100-row limit of the opening"""
    from oracle import motioncrop_oracle as mo

    cases = [
        _video_in_a_frame(96, 64, 1, box=(20, 12, 50, 36)),
        _video_in_a_frame(96, 64, 2, box=(8, 6, 40, 30), second=(56, 30, 30, 28)),
        _video_in_a_frame(96, 64, 3),
        _video_in_a_frame(96, 64, 4, box=(20, 12, 50, 36), stretch=True),
        _video_in_a_frame(160, 112, 5, box=(30, 20, 90, 70)),
        _video_in_a_frame(160, 112, 6, box=(10, 10, 60, 44), second=(84, 56, 64, 48), stretch=True),
        _video_in_a_frame(320, 240, 11, box=(40, 30, 200, 150), second=(250, 170, 60, 56)),  # opening by 10
    ]
    lb = _video_in_a_frame(96, 64, 7, box=(20, 12, 50, 36))
    lb[:, :6, :] = 16  # a letterbox bar on top of it all
    cases.append(lb)
    for k, st in enumerate(cases):
        want = mo.cropdetect_motion(list(st))
        got_h, got_s, got_c = gpu_hash(ctx, st[None], _ffi.CROPDETECT_MOTION)
        assert got_s[0] == 0 and tuple(int(v) for v in got_c[0]) == tuple(want), (k, got_c[0], want)
        l, r, t, b = want
        h, w = st.shape[1:]
        _, want_h, _, _ = o.hash_stack(np.ascontiguousarray(st[:, t:h - b, l:w - r]), 0)
        assert np.array_equal(got_h[0], want_h), k
    assert tuple(int(v) for v in gpu_hash(ctx, cases[0][None], _ffi.CROPDETECT_MOTION)[2][0]) != (0, 0, 0, 0)
    # a batch with different sizes and an error stack in one call
    n = 3
    frames = np.concatenate([cases[0].reshape(-1), cases[4].reshape(-1), cases[2].reshape(-1)])
    d = np.zeros(n, dtype=_ffi.STACK_DESC_DTYPE)
    sizes = [(96, 64), (160, 112), (96, 64)]
    off = 0
    for s, (w, h) in enumerate(sizes):
        d[s] = (off, w * h, w, h, w, 16 if s != 2 else 9, 0, 0)
        off += 16 * w * h
    got_h, got_s, got_c = ctx.hash_stacks(frames, d, _ffi.CROPDETECT_MOTION)
    assert got_s.tolist() == [0, 0, _ffi.STACK_NOT_ENOUGH_FRAMES]
    assert tuple(int(v) for v in got_c[0]) == tuple(mo.cropdetect_motion(list(cases[0])))
    assert tuple(int(v) for v in got_c[1]) == tuple(mo.cropdetect_motion(list(cases[4])))
