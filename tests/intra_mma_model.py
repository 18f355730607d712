"""Lane-exact numpy model of the tensor-core path of intra32_kernel (x266_b200/csrc/intra.cu): reference strip, 4-byte Hankel
windows, the rounding tap on k = 31, mma.m16n8k32 u8 x u8 fragment layouts, the column permutation x = 8(n>>1) + 2t + (n&1) and the
byte-1 extraction.  Driven with the PRODUCT's fragment table (xIntra32MmaTable, pure host code), it must reproduce the checker's
prediction for every fractional angular mode -- a CPU test of the choreography the GPU test then confirms on the device."""
import numpy as np

ANG = [0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21, -26, -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32]
INV = [0] * 11 + [4096, 1638, 910, 630, 482, 390, 315, 256, 315, 390, 482, 630, 910, 1638, 4096] + [0] * 9


def mma_u8(A, B):
    """A[32 lanes][4], B[32][2] (u8x4 registers) -> D[32][4] of m16n8k32 row.col with the PTX fragment layouts"""
    Am = np.zeros((16, 32), np.int64)
    Bm = np.zeros((32, 8), np.int64)
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        for i in range(4):
            Am[g][4 * q + i] = (A[lane][0] >> (8 * i)) & 255
            Am[g + 8][4 * q + i] = (A[lane][1] >> (8 * i)) & 255
            Am[g][16 + 4 * q + i] = (A[lane][2] >> (8 * i)) & 255
            Am[g + 8][16 + 4 * q + i] = (A[lane][3] >> (8 * i)) & 255
            Bm[4 * q + i][g] = (B[lane][0] >> (8 * i)) & 255
            Bm[16 + 4 * q + i][g] = (B[lane][1] >> (8 * i)) & 255
    Dm = Am @ Bm
    return [[int(Dm[l >> 2][2 * (l & 3)]), int(Dm[l >> 2][2 * (l & 3) + 1]), int(Dm[(l >> 2) + 8][2 * (l & 3)]), int(Dm[(l >> 2) + 8][2 * (l & 3) + 1])]
            for l in range(32)]


def window(strip, off):
    return int.from_bytes(bytes(strip[off:off + 4]), "little")


def predict(raw, mode, table, garbage):
    """raw = left[64] | top[65]; table = [35][2][32][4] uint32 from the library; garbage = what the strip held before"""
    strip = bytearray(garbage)
    ver, ang = mode >= 18, ANG[mode]
    ref0 = 36 if ver else 35
    if ver:
        strip[36:100] = bytes(raw[64:128]); strip[ref0 + 64] = raw[128]
    else:
        strip[36:100] = bytes(raw[0:64]); strip[ref0] = raw[64]
    if ang < 0:
        for k in range(1, 33):
            if -k >= ang:
                s = (k * INV[mode] + 128) >> 8
                strip[ref0 - k] = raw[s - 1] if ver else raw[64 + s]
    base = (ang >> 5) + 1 if ang >= 0 else ang + 1
    patch = lambda w, lane: (w & 0x00FFFFFF) | 0x01000000 if (lane & 3) == 3 else w
    T = [[int(table[mode][h][lane][r]) for h in range(2) for r in range(4)] for lane in range(32)]      # 8 words per lane
    P = np.zeros((32, 32), np.uint8)
    for m in range(2):
        D = []
        for t in range(4):
            A, B = [], []
            for lane in range(32):
                g, q = lane >> 2, lane & 3
                if ver:
                    cb = ref0 + base + 4 * q + 8 * (g >> 1) + (g & 1)
                    A.append(T[lane][4 * m:4 * m + 4])
                    B.append([window(strip, cb + 2 * t), patch(window(strip, cb + 16 + 2 * t), lane)])
                else:
                    ca = ref0 + base + 4 * q + g
                    A.append([window(strip, ca + 16 * m), window(strip, ca + 16 * m + 8),
                              patch(window(strip, ca + 16 * m + 16), lane), patch(window(strip, ca + 16 * m + 24), lane)])
                    B.append(T[lane][2 * t:2 * t + 2])
            D.append(mma_u8(A, B))
        for lane in range(32):
            g, q = lane >> 2, lane & 3
            for t in range(4):
                for j in range(4):
                    v = D[t][lane][j]
                    assert 0 <= v < 65536
                    P[16 * m + g + 8 * (j >> 1)][8 * q + 2 * t + (j & 1)] = (v >> 8) & 255
    return P
