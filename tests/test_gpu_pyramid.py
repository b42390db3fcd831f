"""CUDA pyramid (dslam_frame_make_images) vs the oracle's FrameHessian::makeImages restatement: bit-exact."""
import numpy as np
import pytest

import oracle as orc
from direct_stereo_slam_b200 import api
from direct_stereo_slam_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _image(w, h, seed):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 128 + 60 * np.sin(xx * 0.11 + rng.uniform(0, 6)) * np.cos(yy * 0.07) + rng.normal(0, 8, (h, w))
    return np.clip(img, 0, 255).astype(np.float32)


def _compare(fr, dIp_o, abs_o, levels, w, h):
    off = orc.level_offsets(w, h, levels)
    for l in range(levels):
        wl, hl = w >> l, h >> l
        d = fr.dIp(l)
        a = fr.absSquaredGrad(l)
        do = dIp_o[off[l]:off[l + 1]].reshape(hl, wl, 3)
        ao = abs_o[off[l]:off[l + 1]].reshape(hl, wl)
        # intensity: every pixel; gradients: rows 1..h-2 (the reference leaves the first/last row uninitialised)
        assert np.array_equal(d[..., 0].view(np.uint32), do[..., 0].view(np.uint32)), "I level %d" % l
        assert np.array_equal(d[1:-1, :, 1:].view(np.uint32), do[1:-1, :, 1:].view(np.uint32)), "dx/dy level %d" % l
        assert np.array_equal(a[1:-1].view(np.uint32), ao[1:-1].view(np.uint32)), "absSquaredGrad level %d" % l
        assert not d[0, :, 1:].any() and not d[-1, :, 1:].any() and not a[0].any() and not a[-1].any()


@pytest.mark.parametrize("w,h,levels", [(320, 192, 3), (1232, 368, 5), (1024, 768, 5), (154, 46, 2), (77, 23, 1), (150, 94, 2),
                                        (1920, 1200, 5), (1920, 1184, 6)])
def test_make_images_bit_exact(session, oracle, w, h, levels):
    img = _image(w, h, w + h)
    dIp_o, abs_o = oracle.make_images(img, levels)
    fr = api.FrameHessian(session, w, h, levels)
    fr.makeImages(img)
    _compare(fr, dIp_o, abs_o, levels, w, h)
    # a second frame through the same object (buffers are reused) and the download-after-build path
    img2 = _image(w, h, 7)
    dIp_o, abs_o = oracle.make_images(img2, levels)
    fr.upload(img2)
    fr.build()
    fr.download()
    _compare(fr, dIp_o, abs_o, levels, w, h)
    fr.close()


def test_make_images_gamma_weight(session, oracle):
    """HCalib != 0 && setting_gammaWeightsPixelSelect == 1: absSquaredGrad *= (B[c+1]-B[c])^2."""
    w, h, levels = 640, 480, 5
    img = _image(w, h, 11)
    B = (255.0 * (np.arange(256) / 255.0) ** 0.8).astype(np.float32)
    dIp_o, abs_o = oracle.make_images(img, levels, B256=B)
    fr = api.FrameHessian(session, w, h, levels)
    fr.makeImages(img, B256=B)
    _compare(fr, dIp_o, abs_o, levels, w, h)
    fr.close()


def test_non_finite_input(session, oracle):
    """non-finite gradients are zeroed (HessianBlocks.cpp:174-177)."""
    w, h, levels = 128, 96, 2
    img = _image(w, h, 5)
    img[40, 50] = np.inf
    img[10, 0] = np.nan
    dIp_o, abs_o = oracle.make_images(img, levels)
    fr = api.FrameHessian(session, w, h, levels)
    fr.makeImages(img)
    off = orc.level_offsets(w, h, levels)
    for l in range(levels):
        wl, hl = w >> l, h >> l
        d = fr.dIp(l)[1:-1]
        do = dIp_o[off[l]:off[l + 1]].reshape(hl, wl, 3)[1:-1]
        assert np.array_equal(np.isnan(d), np.isnan(do))
        m = ~np.isnan(do)
        assert np.array_equal(d[m], do[m])
    fr.close()


def test_synthetic_scene_frames(session, oracle):
    c = syn.make_tracking_case("tiny", 2)
    cfg = c["cfg"]
    levels = orc.pyr_levels_used(cfg["w"], cfg["h"])
    for key in ("img_ref", "img_new", "img_right"):
        dIp_o, abs_o = oracle.make_images(c[key], levels)
        fr = api.FrameHessian(session, cfg["w"], cfg["h"], levels)
        fr.makeImages(c[key])
        _compare(fr, dIp_o, abs_o, levels, cfg["w"], cfg["h"])
        fr.close()


def test_upload_batch_equals_single_uploads(session, oracle):
    """dslam_frame_upload_batch: images that lie back to back in host memory travel as one transfer (+ a scatter kernel),
    stragglers one by one — the pyramids are the same bits either way (and equal the oracle's)."""
    w, h, levels = 320, 192, 3
    arena = session.pinned((5, h, w))
    for k in range(5):
        arena[k] = _image(w, h, 40 + k)
    lone = _image(w, h, 99)                                  # not part of the arena
    imgs = [arena[0], arena[1], arena[2], lone, arena[4], arena[3]]  # a run of 3, a straggler, two out-of-order slots
    frames = [api.FrameHessian(session, w, h, levels) for _ in imgs]
    api.upload_frames(frames, imgs)
    api.build_frames(frames)
    for fr, im in zip(frames, imgs):
        fr.download()
        dIp_o, abs_o = oracle.make_images(np.array(im), levels)
        _compare(fr, dIp_o, abs_o, levels, w, h)
        fr.close()


def test_async_build_overlaps_tracking_and_stays_ordered(session, oracle):
    """stage_host bit 2 (build on the pyramid stream): the pyramid, its host mirror and a tracking call on the frame are the
    same bits as with a synchronous build; a tracking call on ANOTHER frame queued after the asynchronous build is unaffected;
    re-uploading a frame whose asynchronous build may still be in flight is ordered behind it."""
    from helpers import IDENT7, GpuCase, OracleCase

    oc = OracleCase(oracle, "tiny", 5)
    gc = GpuCase(session, oc, template="device")
    c = oc.case
    ref = gc.trk.trackNewestCoarse(gc.f_new, IDENT7, (0, 0), oc.levels - 1)          # synchronous baseline
    levels, w, h = oc.levels, oc.w, oc.h
    f2 = api.FrameHessian(session, w, h, levels)
    for rep in range(3):
        f2.upload(c["img_right"] if rep == 1 else c["img_new"])
        api.build_frames([f2], stage_host=3, overlap=True)
        other = gc.trk.trackNewestCoarse(gc.f_new, IDENT7, (0, 0), oc.levels - 1)    # other frame: runs beside the build
        assert other[0] == ref[0] and np.array_equal(other[1], ref[1])
        f2.download()
        dIp_o, abs_o = oracle.make_images(c["img_right"] if rep == 1 else c["img_new"], levels)
        _compare(f2, dIp_o, abs_o, levels, w, h)
    got = gc.trk.trackNewestCoarse(f2, IDENT7, (0, 0), oc.levels - 1)                # the asynchronously built frame itself
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
    f2.close()
    gc.close()


def test_download_of_an_array_the_current_build_did_not_stage(session):
    """A frame object is rebuilt with only ONE host-layout copy staged while the staging buffer of the other one still holds the
    previous image: a later download of that other array must unpack the CURRENT pyramid, not hand out the stale buffer."""
    rng = np.random.default_rng(5)
    w, h, levels = 320, 192, 3
    img_a = np.clip(rng.normal(128, 40, (h, w)), 0, 255).astype(np.float32)
    img_b = np.clip(rng.normal(90, 30, (h, w)), 0, 255).astype(np.float32)
    want = api.FrameHessian(session, w, h, levels)
    want.makeImages(img_b)
    f = api.FrameHessian(session, w, h, levels)
    f.upload(img_a)
    api.build_frames([f], stage_host=3)
    f.download()                              # both staging buffers now hold image A
    f.upload(img_b)
    api.build_frames([f], stage_host=1)       # image B: only dIp staged
    f.download()                              # dIp from the staging copy, absSquaredGrad unpacked from the texels
    assert np.array_equal(f.dIp_all.view(np.uint32), want.dIp_all.view(np.uint32))
    assert np.array_equal(f.absSquaredGrad_all.view(np.uint32), want.absSquaredGrad_all.view(np.uint32))
    f.upload(img_a)
    api.build_frames([f], stage_host=2)       # the other way round
    f.download()
    want.makeImages(img_a)
    assert np.array_equal(f.dIp_all.view(np.uint32), want.dIp_all.view(np.uint32))
    assert np.array_equal(f.absSquaredGrad_all.view(np.uint32), want.absSquaredGrad_all.view(np.uint32))
    f.close()
    want.close()
