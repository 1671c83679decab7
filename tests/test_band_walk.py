"""CPU: the item walk of the persistent bev_band CTAs (csrc/bev_rasterize.cu: band_cta_count, `band_rot`, the (f, band) increments
of bev_band_kernel), restated in Python: whatever the launch shape, every (ring frame, band) item is visited exactly once, and
with the rotation a CTA no longer meets the same band in every frame it visits."""
import pytest

SLOTS = 4 * 148


def cta_count(n_items, slots=SLOTS):
    if n_items <= slots:
        return n_items
    per_cta = -(-n_items // slots)
    return -(-n_items // per_cta)


def walk(n_items, nb, G):
    """Returns {cta: [(f, band), ...]} exactly as the kernel steps: host-side step_f / step_b / band_rot, device-side increments."""
    step_f, step_b = G // nb, G % nb
    rot = nb // 2 if (G < n_items and G % nb == 0) else 0
    visits = {}
    for c in range(G):
        item, f, band = c, 0, c
        while band >= nb:
            band -= nb
            f += 1
        seq = []
        while True:
            seq.append((f, band))
            item_next = item + G
            if item_next >= n_items:
                break
            f_next, band_next = f + step_f, band + step_b
            if band_next >= nb:
                band_next -= nb
                f_next += 1
            band_next += rot
            if band_next >= nb:
                band_next -= nb
            item, f, band = item_next, f_next, band_next
        visits[c] = seq
    return visits


@pytest.mark.parametrize("nb,frames", [(128, 8), (128, 16), (128, 1), (128, 5), (128, 64), (100, 8), (37, 50), (160, 12), (128, 3)])
def test_every_item_exactly_once(nb, frames):
    n_items = nb * frames
    for G in {cta_count(n_items), min(n_items, SLOTS), min(n_items, 3 * 148)}:
        seen = [fb for seq in walk(n_items, nb, G).values() for fb in seq]
        assert len(seen) == n_items and len(set(seen)) == n_items
        assert all(0 <= f < frames and 0 <= b < nb for f, b in seen)


def test_kitti_chunk_shape_and_rotation():
    n_items, nb = 8 * 128, 128
    G = cta_count(n_items)
    assert G == 512                                  # two items on every CTA instead of 432 x 2 + 160 x 1
    for seq in walk(n_items, nb, G).values():
        assert len(seq) == 2 and seq[0][1] != seq[1][1] and (seq[1][1] - seq[0][1]) % nb == nb // 2
