"""CPU model of the tile geometry and tap coordinates in experiments/conv_stride2_tma.patch.

The TMA box load is modelled with the semantics measured by tma_stride2_probe.cu (box dims count source elements, every
`stride`-th element is loaded and compacted, coordinates outside the tensor read as zero).  With that model, the A tiles the
patched vn_gemm would fetch - tile geometry over the OUTPUT grid, start coordinate (w0*cs + dx - cpad, h0*cs + dy - cpad),
k = tap*C + c - are multiplied with the weights and compared with F.conv2d at stride 2 for both padding forms, including
ragged edges and images smaller than a tile.   python experiments/stride2_coords_model.py
"""
import torch
import torch.nn.functional as F

BM = 128


def tma_box(x, w_start, h_start, img, box_w, box_h, es):
    """x [nb,H,W,C]; returns the compacted [ceil(box_h/es) * ceil(box_w/es), C] rows of one box load."""
    nb, H, W, C = x.shape
    rows = []
    for hh in range(h_start, h_start + box_h, es):
        for ww in range(w_start, w_start + box_w, es):
            ok = 0 <= hh < H and 0 <= ww < W and 0 <= img < nb
            rows.append(x[img, hh, ww] if ok else torch.zeros(C, dtype=x.dtype))
    return torch.stack(rows)


def conv_via_tiles(x, wk, cs, cpad):
    nb, H, W, C = x.shape
    N = wk.shape[0]
    Ho = H if cs == 1 else ((H - 1) // 2 + 1 if cpad else (H - 2) // 2 + 1)
    Wo = W if cs == 1 else ((W - 1) // 2 + 1 if cpad else (W - 2) // 2 + 1)
    tw = 1
    while tw < 64 and Wo % (tw * 2) == 0:
        tw *= 2
    th = BM // tw
    while th > 1 and th // 2 >= Ho:
        th //= 2
    assert cs * tw <= 256 and cs * th <= 256
    tiles_w, tiles_h = -(-Wo // tw), -(-Ho // th)
    out = torch.zeros(nb, Ho, Wo, N, dtype=torch.float64)
    for img in range(nb):
        for t in range(tiles_w * tiles_h):
            h0, w0 = (t // tiles_w) * th, (t % tiles_w) * tw
            ah0, aw0 = h0 * cs - cpad, w0 * cs - cpad
            acc = torch.zeros(tw * th, N, dtype=torch.float64)
            for tap in range(9):
                dy, dx = tap // 3, tap % 3
                a = tma_box(x, aw0 + dx, ah0 + dy, img, cs * tw, cs * th, cs)          # [tw*th, C]
                assert a.shape[0] == tw * th
                acc += a.double() @ wk[:, tap * C:(tap + 1) * C].double().t()
            for r in range(tw * th):                                                   # tile_row(): clip to the output grid
                h, w = h0 + r // tw, w0 + r % tw
                if h < Ho and w < Wo:
                    out[img, h, w] = acc[r]
    return out


def main():
    g = torch.Generator().manual_seed(0)
    worst = 0.0
    for (nb, H, W, C, N, cs, cpad) in [(1, 8, 8, 4, 3, 2, 1), (2, 9, 13, 4, 5, 2, 1), (1, 64, 64, 2, 2, 2, 1), (1, 6, 10, 4, 3, 2, 0),
                                       (1, 16, 24, 2, 3, 2, 0), (2, 5, 7, 3, 2, 2, 1), (1, 130, 6, 2, 2, 2, 0), (1, 12, 12, 4, 3, 1, 1),
                                       (1, 2, 2, 3, 2, 2, 0), (1, 1, 1, 3, 2, 2, 1)]:
        x = torch.randn(nb, H, W, C, generator=g)
        w4 = torch.randn(N, C, 3, 3, generator=g)
        wk = w4.permute(0, 2, 3, 1).reshape(N, 9 * C)
        got = conv_via_tiles(x, wk, cs, cpad)
        xi = x.permute(0, 3, 1, 2).double()
        if cs == 2 and cpad == 0:
            ref = F.conv2d(F.pad(xi, (0, 1, 0, 1)), w4.double(), stride=2)
        else:
            ref = F.conv2d(xi, w4.double(), stride=cs, padding=1)
        ref = ref.permute(0, 2, 3, 1)
        assert got.shape == ref.shape, (got.shape, ref.shape)
        err = float((got - ref).abs().max())
        worst = max(worst, err)
        print(f"nb{nb} {H}x{W} C{C} N{N} stride {cs} pad {cpad}: out {tuple(ref.shape[1:3])}, max |diff| {err:.2e}")
    assert worst < 1e-10
    print("OK")


if __name__ == "__main__":
    main()
