"""Host-side restatement of the dropout masks the CUDA kernels draw (common.cuh: philox4x32 / drop_keep1 / drop_keep8 /
drop_bits_c), as dense keep-scale tensors in the oracle's layouts.  Test infrastructure only: it lets the parity tests
feed the SAME masks to `oracle.qavit_oracle.quad_block(..., masks=...)` and compare a train-mode dropout run of the
CUDA path element by element instead of statistically.

Site ids and element-id schemes mirror qa-vit_b200/csrc/block.cu (enum DS_*), attn.cu / cga.cu (SIMT, fp32 runs),
attn_mma.cu / attn_msda64.cu / cga_mma.cu / cga_mma64.cu (mma.sync C-fragment layout, bf16 runs) and drop.cu."""
import numpy as np
import torch

DS_ATT, DS_PROJ, DS_B1, DS_B2, DS_FFN, DS_PATH = 1, 5, 9, 10, 11, 12
M32 = np.uint64(0xFFFFFFFF)
PHILOX_ROUNDS = 7


def philox4x32(key0, key1, c0, c1, c2, c3):
    """Philox4x32-7 (QV_PHILOX_ROUNDS in common.cuh) on uint64 numpy arrays holding 32-bit values (broadcastable)."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & M32 for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint64(key0 & 0xFFFFFFFF), np.uint64(key1 & 0xFFFFFFFF)
    for _ in range(PHILOX_ROUNDS):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & M32
        hi1, lo1 = p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M32, lo1, (hi0 ^ c3 ^ k1) & M32, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    return c0, c1, c2, c3


class Site:
    def __init__(self, seed: int, offset: int, site: int, p: float):
        self.k0 = seed & 0xFFFFFFFF
        self.k1 = ((seed >> 32) & 0xFFFFFFFF) ^ site
        self.off_lo, self.off_hi = offset & 0xFFFFFFFF, (offset >> 32) & 0xFFFFFFFF
        self.thr = int(np.float32(np.float32(p) * np.float32(65536.0)) + np.float32(0.5))
        self.inv = float(np.float32(65536.0) / np.float32(65536.0 - self.thr))

    def _scale(self, r16):
        return torch.from_numpy(np.where(r16 >= self.thr, np.float32(self.inv), np.float32(0.0)).astype(np.float32))

    def keep1(self, ids):
        """drop_keep1: one element per Philox call."""
        ids = np.asarray(ids, dtype=np.uint64)
        r = philox4x32(self.k0, self.k1, ids & M32, ids >> np.uint64(32), self.off_lo, self.off_hi ^ 0x51)
        return self._scale(r[0] & np.uint64(0xFFFF))

    def keep_rows(self, rows: int, C: int):
        """drop_rows on a contiguous [rows, C] matrix: drop_keep8 per 8 consecutive elements."""
        n8 = rows * C // 8
        ids = np.arange(n8, dtype=np.uint64)
        r = philox4x32(self.k0, self.k1, ids & M32, ids >> np.uint64(32), self.off_lo, self.off_hi ^ 0x58)
        w = np.stack(r, axis=1)                                              # [n8, 4]
        r16 = np.stack([w & np.uint64(0xFFFF), w >> np.uint64(16)], axis=2)  # [n8, 4, 2] -> element 2 i + half
        return self._scale(r16.reshape(rows, C))

    def keep_tiles(self, tiles, ncols: int):
        """drop_bits_c: [len(tiles), 16, ncols] keep scales of 16-row tiles held in mma C-fragment layout."""
        tiles = np.asarray(tiles, dtype=np.uint64)
        NT = ncols // 8
        i = np.arange(16)[:, None]
        j = np.arange(ncols)[None, :]
        lane = (i % 8) * 4 + (j % 8) // 2
        k = 4 * (j // 8) + (j % 2) + 2 * (i // 8)                            # bit index 4 n + e
        call, word, half = k >> 3, (k & 7) >> 1, k & 1
        out = np.zeros((len(tiles), 16, ncols), dtype=np.uint64)
        for c in range((NT + 1) // 2):
            lanes = np.arange(32, dtype=np.uint64)
            r = philox4x32(self.k0, self.k1, tiles[:, None], (lanes | np.uint64(c << 5))[None, :], self.off_lo, self.off_hi ^ 0x5C)
            w = np.stack(r, axis=2)                                          # [T, 32, 4]
            sel = call == c
            vals = w[:, lane, word]                                          # [T, 16, ncols]
            vals = np.where(half == 1, vals >> np.uint64(16), vals & np.uint64(0xFFFF))
            out = np.where(sel[None], vals, out)
        return self._scale(out)


def block_masks(seed: int, offset: int, p: float, p_path: float, B: int, Nt: int, *, bf16: bool, d: int = 192, H: int = 4,
                G: int = 6, kb: int = 16, klin: int = 32, ws: int = 4, bh: int = 96):
    """Keep-scale tensors of one quad-block forward call (rng snapshot {seed, offset}) in the oracle's layouts."""
    side = int(round(Nt ** 0.5))
    R = B * Nt
    nkv = klin + kb
    m = {}
    S = lambda site: Site(seed, offset, site, p)
    if p > 0:
        rows = np.arange(R, dtype=np.uint64)
        if not bf16:      # SIMT kernels: id = ((row * H + h) << 7) | key
            def simt(site, NKV):
                ids = ((rows[:, None, None] * np.uint64(H) + np.arange(H, dtype=np.uint64)[None, :, None]) << np.uint64(7)) | \
                    np.arange(NKV, dtype=np.uint64)[None, None, :]
                return S(site).keep1(ids).reshape(B, Nt, H, NKV).permute(0, 2, 1, 3)      # [B, H, Nt, NKV]
            att_swa_img, m["att_msda"], m["att_cross"] = simt(DS_ATT + 0, nkv), simt(DS_ATT + 1, nkv), simt(DS_ATT + 3, kb)
            NKVc = Nt + kb
            ids = ((((rows[:, None, None, None] * np.uint64(G) + np.arange(G, dtype=np.uint64)[None, :, None, None]) * np.uint64(4)
                     + np.arange(4, dtype=np.uint64)[None, None, :, None]) << np.uint64(7))
                   | np.arange(NKVc, dtype=np.uint64)[None, None, None, :])
            m["att_cga"] = S(DS_ATT + 2).keep1(ids).reshape(B, Nt, G, 4, NKVc).permute(0, 2, 3, 1, 4).reshape(B * G, 4, Nt, NKVc)
        else:             # mma.sync kernels: 16-row tiles in C-fragment layout
            nws = side // ws
            nW = nws * nws
            assert ws * ws == 16
            # SWA (attn_mma.cu): tile = (b * nW + wi) * H + h, rows = in-window index -> directly the oracle's window layout
            t = S(DS_ATT + 0).keep_tiles(np.arange(B * nW * H), nkv)
            m["att_swa"] = t.reshape(B * nW, H, 16, nkv)
            att_swa_img = None
            T = Nt // 16
            if Nt == 16:  # attn_mma.cu, tile = b * H + h
                m["att_msda"] = S(DS_ATT + 1).keep_tiles(np.arange(B * H), nkv).reshape(B, H, 16, nkv)
            else:         # attn_msda64.cu, tile = (b * H + h) * T + qt
                m["att_msda"] = S(DS_ATT + 1).keep_tiles(np.arange(B * H * T), nkv).reshape(B, H, Nt, nkv)
            # cross (attn_mma.cu after as_tiles16): tile = (b * T + qt) * H + h
            t = S(DS_ATT + 3).keep_tiles(np.arange(B * T * H), kb).reshape(B, T, H, 16, kb)
            m["att_cross"] = t.permute(0, 2, 1, 3, 4).reshape(B, H, Nt, kb)
            NKVc = Nt + kb
            if Nt == 16:  # cga_mma.cu: tile = (b * G + grp) * NH + h
                m["att_cga"] = S(DS_ATT + 2).keep_tiles(np.arange(B * G * 4), NKVc).reshape(B * G, 4, 16, NKVc)
            else:         # cga_mma64.cu: tile = ((b * T + qt) * G + grp) * NH + h
                t = S(DS_ATT + 2).keep_tiles(np.arange(B * T * G * 4), NKVc).reshape(B, T, G, 4, 16, NKVc)
                m["att_cga"] = t.permute(0, 2, 3, 1, 4, 5).reshape(B * G, 4, Nt, NKVc)
        if att_swa_img is not None:   # image order [B, H, Nt, NKV] -> the oracle's window partition [B * nW, H, ws * ws, NKV]
            nws = side // ws
            t = att_swa_img.reshape(B, H, nws, ws, nws, ws, nkv).permute(0, 2, 4, 1, 3, 5, 6)
            m["att_swa"] = t.reshape(B * nws * nws, H, ws * ws, nkv)
        for i, name in enumerate(("swa", "msda", "cga", "cross")):
            m["proj_" + name] = S(DS_PROJ + i).keep_rows(R, d).reshape(B, Nt, d)
        m["b1"] = S(DS_B1).keep_rows(R, bh).reshape(B, Nt, bh)
        m["b2"] = S(DS_B2).keep_rows(R, d).reshape(B, Nt, d)
        m["ffn"] = S(DS_FFN).keep_rows(R, d).reshape(B, Nt, d)
    if p_path > 0:
        rs = Site(seed, offset, DS_PATH, p_path).keep1(np.arange(2 * B, dtype=np.uint64))
        m["path1"], m["path2"] = rs[:B].reshape(B, 1, 1), rs[B:].reshape(B, 1, 1)
    return m
