"""STOM placement policy (model/STOM.py:72-141 after the tracker) on the CPU: the oracle restatement against the
golden frames produced by the reference's own ``propagate_in_video`` (tests/golden/make_golden.py, tracker stubbed),
its numpy/OpenCV building blocks against the live libraries, and the host policy ``stom_frame_ops`` against both."""
import os

import numpy as np
import pytest

from oracle import overlay_ref, stom_policy_ref as sp

CASES = ["rect_a", "rect_b", "mask_odd", "mask_even", "mask_small"]


def _case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "stom_policy.npz"))
    fr, layer, tr, vi, meta, out = [g[f"{name}_{k}"] for k in ("frames", "layer", "tracks", "vis", "meta", "out")]
    return fr, layer, tr, vi, int(meta[0]), ("mask" if meta[1] else "rectangle"), out


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_propagate_in_video(golden_dir, name):
    fr, layer, tr, vi, key, shape, out = _case(golden_dir, name)
    assert np.array_equal(sp.propagate_ref(fr, layer, tr, vi, key, shape), out)


def test_fixture_covers_every_policy_branch(golden_dir):
    """The golden clips must exercise: shifted layer, circle stamp, too-few-visible, none-visible, NaN track."""
    seen = set()
    for name in CASES:
        fr, layer, tr, vi, key, shape, out = _case(golden_dir, name)
        for i in range(len(fr)):
            changed = bool((out[i] != fr[i]).any())
            if i == key:
                seen.add("key")
            elif not vi[i].any():
                assert not changed
                seen.add("none_visible")
            elif np.isnan(tr[i][vi[i]]).any():
                assert not changed
                seen.add("nan")
            elif vi[i].sum() < vi.shape[1] // 2:
                assert not changed
                seen.add("few_visible")
            elif changed:
                seen.add("circle" if shape == "mask" else "shift")
    assert seen == {"key", "none_visible", "nan", "few_visible", "circle", "shift"}


def test_numpy_float32_reductions_bit_exact():
    rng = np.random.default_rng(3)
    for n in list(range(1, 40)) + [127, 128, 129, 255, 256, 257, 1000, 4097, 9999]:
        a = (rng.standard_normal((n, 2)) * 100).astype(np.float32)
        col = a[:, 1]                                                  # strided, as filtered_flows[:, 1] (:126)
        assert np.add.reduce(col).tobytes() == sp.pairwise_sum_f32(col).tobytes()
        assert np.mean(col).tobytes() == sp.mean_f32(col).tobytes()
        assert np.float32(np.median(np.abs(col))).tobytes() == sp.median_f32(np.abs(col)).tobytes()
        mag = np.sqrt((a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1]).astype(np.float32))
        assert np.linalg.norm(a, axis=1).tobytes() == mag.tobytes()


def test_opencv_structuring_element_and_closing():
    cv2 = pytest.importorskip("cv2")
    for k in range(1, 64):
        se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
        mine = np.zeros((k, k), np.uint8)
        for i, (j1, j2) in enumerate(sp.ellipse_rows_ref(k)):
            mine[i, j1:j2] = 1
        assert np.array_equal(se, mine), k
    rng = np.random.default_rng(0)
    for h, w, k in [(60, 90, 4), (75, 75, 5), (150, 135, 9), (150, 180, 10), (64, 64, 1), (80, 70, 2), (33, 47, 6)]:
        for dens in (1, 3, 30, 300):
            m = np.zeros((h, w), np.uint8)
            m[rng.integers(0, h, dens), rng.integers(0, w, dens)] = 255
            ref = cv2.morphologyEx(m, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k)))
            assert np.array_equal(sp.close_ref(m, k), ref), (h, w, k, dens)


def _apply_ops(frames, layer, ops):
    """Composite a clip from FrameOp records with the overlay oracle."""
    per = []
    for o in ops:
        if o.mode == 1:
            per.append(dict(mode=1, sx=o.sx, zx=o.zx, sy=o.sy, zy=o.zy))
        elif o.mode == 2:
            per.append(dict(mode=2, cx=o.cx, cy=o.cy, r=o.r, rgba=tuple(o.rgba)))
        else:
            per.append(dict(mode=0))
    return overlay_ref.overlay_clip_ref(frames, layer, per)


@pytest.mark.parametrize("name", CASES)
def test_host_policy_matches_reference(golden_dir, name):
    """rga3_release_b200.stom_frame_ops (numpy + cv2 on the host) -> ops -> oracle composite == reference frames."""
    pytest.importorskip("cv2")
    import rga3_release_b200 as vit
    fr, layer, tr, vi, key, shape, out = _case(golden_dir, name)
    h, w = layer.shape[:2]
    ops = vit.stom_frame_ops(tr, vi, key, shape, h, w, layer)
    assert np.array_equal(_apply_ops(fr, layer, ops), out)


def test_shift_closed_form_matches_host_shift_from_flow():
    """The closed form the device kernel uses (csrc/stom_policy.cu::shift_from_flow) == overlay.shift_from_flow."""
    import rga3_release_b200 as vit
    rng = np.random.default_rng(5)
    flows = np.concatenate([rng.uniform(-70, 70, 400), np.arange(-5, 6), [-0.5, -63.99, -64.0, -64.01, 63.5, -0.0, 1e-7,
                                                                         -1e-7]]).astype(np.float32)
    for n in (1, 2, 28, 64):
        for f32 in flows:
            f = float(f32)
            fl = np.floor(f)
            z = int(f < 0 and f != fl and np.floor(-f) < n)
            assert (int(fl), z) == vit.shift_from_flow(f, n), (f, n)
