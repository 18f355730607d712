"""Lane-exact numpy model of x266_b200/csrc/dct_imma.cu: the PTX m16n8k32 fragment layouts plus the
kernel's index permutations (sigma, pi, pi2), byte-plane split and PRMT selectors.  It lets the CPU
suite prove the register choreography of the tensor-core kernel against the oracle without a GPU."""
import numpy as np


def prmt(a, b, sel):
    src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    out = 0
    for i in range(4):
        out |= src[(sel >> (4 * i)) & 0x7] << (8 * i)
    return out


def s8(x):
    x &= 0xFF
    return x - 256 if x >= 128 else x


def mma_m16n8k32(A_frag, B_frag, C_frag, b_signed):
    """A_frag[lane][4], B_frag[lane][2] (u32 regs), C_frag[lane][4] (ints) -> D_frag[lane][4].
    Layouts per PTX ISA 'mma.m16n8k32' (8-bit): groupID = lane>>2, tig = lane&3."""
    A = np.zeros((16, 32), np.int64)
    B = np.zeros((32, 8), np.int64)
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        for r in range(4):
            row = g + 8 * (r & 1)
            for i in range(4):
                A[row, 16 * (r >> 1) + 4 * q + i] = s8(A_frag[lane][r] >> (8 * i))
        for r in range(2):
            for i in range(4):
                v = (B_frag[lane][r] >> (8 * i)) & 0xFF
                B[16 * r + 4 * q + i, g] = s8(v) if b_signed else v
    D = A @ B
    out = []
    for lane in range(32):
        g, q = lane >> 2, lane & 3
        out.append([int(D[g + 8 * (c >> 1), 2 * q + (c & 1)]) + C_frag[lane][c] for c in range(4)])
    return out


def perm_sigma(mu):
    return 8 * ((mu >> 1) & 3) + 2 * (mu >> 3) + (mu & 1)


def perm_pi(k):
    return 8 * ((k & 15) >> 2) + 4 * (k >> 4) + (k & 3)


def perm_pi2(k):
    hi, q, i = k >> 4, (k >> 2) & 3, k & 3
    return 8 * (2 * hi + (i >> 1)) + 2 * q + (i & 1)


def pack4(vals):
    return sum((int(v) & 0xFF) << (8 * i) for i, v in enumerate(vals))


def dct32_imma_model(block, g32, s1, s2):
    """block: [32,32] int16 -> [32,32] int16 exactly as the kernel's lanes would produce it."""
    raw = block.astype(np.int16).tobytes()
    words = np.frombuffer(raw, np.uint32)          # 512 words of the 2 KiB block
    A1 = [[[0] * 4 for _ in range(32)] for _ in range(2)]
    A2 = [[[0] * 4 for _ in range(32)] for _ in range(2)]
    for m in range(2):
        for lane in range(32):
            g, q = lane >> 2, lane & 3
            for r in range(4):
                row = 16 * m + g + 8 * (r & 1)
                kb = 16 * (r >> 1) + 4 * q
                k1 = perm_sigma(row)
                A1[m][lane][r] = pack4([g32[k1][perm_pi(kb + i)] for i in range(4)])
                A2[m][lane][r] = pack4([g32[row][perm_pi2(kb + i)] for i in range(4)])
    BL = [[[0, 0] for _ in range(32)] for _ in range(4)]
    BH = [[[0, 0] for _ in range(32)] for _ in range(4)]
    for t in range(4):
        for lane in range(32):
            w = [int(words[(t * 512 + lane * 16) // 4 + i]) for i in range(4)]
            BL[t][lane] = [prmt(w[0], w[1], 0x6420), prmt(w[2], w[3], 0x6420)]
            BH[t][lane] = [prmt(w[0], w[1], 0x7531), prmt(w[2], w[3], 0x7531)]
    add1, add2 = 1 << (s1 - 1), 1 << (s2 - 1)
    cA1 = [[add1] * 4 for _ in range(32)]
    cA2 = [[add2] * 4 for _ in range(32)]
    cZ = [[0] * 4 for _ in range(32)]
    B2L = [[[0, 0] for _ in range(32)] for _ in range(4)]
    B2H = [[[0, 0] for _ in range(32)] for _ in range(4)]
    for m in range(2):
        r = [[[0] * 4 for _ in range(32)] for _ in range(4)]
        for t in range(4):
            dl = mma_m16n8k32(A1[m], BL[t], cA1, False)
            dh = mma_m16n8k32(A1[m], BH[t], cZ, True)
            for lane in range(32):
                for c in range(4):
                    r[t][lane][c] = ((dl[lane][c] + dh[lane][c] * 256) >> s1) & 0xFFFFFFFF
        for h in range(2):
            for lane in range(32):
                p = [prmt(r[t][lane][2 * h], r[t][lane][2 * h + 1], 0x5140) for t in range(4)]
                B2L[2 * m + h][lane] = [prmt(p[0], p[1], 0x5410), prmt(p[2], p[3], 0x5410)]
                B2H[2 * m + h][lane] = [prmt(p[0], p[1], 0x7632), prmt(p[2], p[3], 0x7632)]
    out = np.zeros(512, np.uint32)
    for m2 in range(2):
        r = [[[0] * 4 for _ in range(32)] for _ in range(4)]
        for t2 in range(4):
            dl = mma_m16n8k32(A2[m2], B2L[t2], cA2, False)
            dh = mma_m16n8k32(A2[m2], B2H[t2], cZ, True)
            for lane in range(32):
                for c in range(4):
                    r[t2][lane][c] = ((dl[lane][c] + dh[lane][c] * 256) >> s2) & 0xFFFFFFFF
        for h in range(2):
            for lane in range(32):
                g, q = lane >> 2, lane & 3
                o = [prmt(r[t][lane][2 * h], r[t][lane][2 * h + 1], 0x5410) for t in range(4)]
                elem = (16 * m2 + 8 * h + g) * 32 + q * 8         # int16 index of the 128-bit store
                for i in range(4):
                    out[elem // 2 + i] = o[i]
    return np.frombuffer(out.tobytes(), np.int16).reshape(32, 32).copy()
