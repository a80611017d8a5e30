"""BC7 block decoder (test infrastructure): the format's decoding rules for every mode that stores opaque colour with two or
one subsets without rotation (modes 1, 3, 6 — the ones ConvertToBinary's encoder profile emits), written from the format
specification, not from either encoder.  Used to compare the decoded error of two encodings of the same image."""
import numpy as np

SUBSET1_MASKS = [
    0xCCCC, 0x8888, 0xEEEE, 0xECC8, 0xC880, 0xFEEC, 0xFEC8, 0xEC80, 0xC800, 0xFFEC, 0xFE80, 0xE800, 0xFFE8, 0xFF00, 0xFFF0,
    0xF000, 0xF710, 0x008E, 0x7100, 0x08CE, 0x008C, 0x7310, 0x3100, 0x8CCE, 0x088C, 0x3110, 0x6666, 0x366C, 0x17E8, 0x0FF0,
    0x718E, 0x399C, 0xAAAA, 0xF0F0, 0x5A5A, 0x33CC, 0x3C3C, 0x55AA, 0x9696, 0xA55A, 0x73CE, 0x13C8, 0x324C, 0x3BDC, 0x6996,
    0xC33C, 0x9966, 0x0660, 0x0272, 0x04E4, 0x4E40, 0x2720, 0xC936, 0x936C, 0x39C6, 0x639C, 0x9336, 0x9CC6, 0x817E, 0xE718,
    0xCCF0, 0x0FCC, 0x7744, 0xEE22]
ANCHORS = [15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 2, 8, 2, 2, 8, 8, 15, 2, 8, 2, 2, 8, 8, 2, 2,
           15, 15, 6, 8, 2, 8, 15, 15, 2, 8, 2, 2, 2, 15, 15, 6, 6, 2, 6, 8, 15, 15, 2, 2, 15, 15, 15, 15, 15, 2, 2, 15]
WEIGHTS = {2: [0, 21, 43, 64], 3: [0, 9, 18, 27, 37, 46, 55, 64],
           4: [0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64]}


class _Bits:
    def __init__(self, block):
        self.v = int.from_bytes(bytes(block), "little")
        self.pos = 0

    def get(self, n):
        r = (self.v >> self.pos) & ((1 << n) - 1)
        self.pos += n
        return r


def block_mode(block):
    """Mode number = position of the lowest set bit of the first byte (8 = invalid)."""
    b = int(block[0])
    return (b & -b).bit_length() - 1 if b else 8


def decode_block(block):
    """16 bytes -> uint8 [16, 4] (pixel y * 4 + x; R, G, B, A).  Modes 1, 3 and 6."""
    mode = block_mode(block)
    r = _Bits(block)
    r.get(mode + 1)
    if mode == 6:
        ends = np.zeros((2, 4), np.int64)
        for c in range(4):
            for e in range(2):
                ends[e, c] = r.get(7)
        for e in range(2):
            ends[e] = ends[e] * 2 + r.get(1)
        idx = [r.get(3 if k == 0 else 4) for k in range(16)]
        assert r.pos == 128
        w = np.array([WEIGHTS[4][i] for i in idx])[:, None]
        return ((ends[0][None] * (64 - w) + ends[1][None] * w + 32) >> 6).astype(np.uint8)
    if mode not in (1, 3):
        raise ValueError("mode %d is not emitted by the veryfast opaque profile" % mode)
    part = r.get(6)
    ebits, ibits = (6, 3) if mode == 1 else (7, 2)
    ends = np.zeros((4, 3), np.int64)  # subset 0 end 0, subset 0 end 1, subset 1 end 0, subset 1 end 1
    for c in range(3):
        for e in range(4):
            ends[e, c] = r.get(ebits)
    if mode == 1:  # one parity bit per subset, then 7 -> 8 bits
        for s in range(2):
            p = r.get(1)
            for e in (2 * s, 2 * s + 1):
                v = ends[e] * 2 + p
                ends[e] = (v << 1) | (v >> 6)
    else:  # one parity bit per end point
        for e in range(4):
            ends[e] = ends[e] * 2 + r.get(1)
    anchor = ANCHORS[part]
    idx = [r.get(ibits - 1 if k in (0, anchor) else ibits) for k in range(16)]
    assert r.pos == 128
    out = np.zeros((16, 4), np.uint8)
    out[:, 3] = 255
    for k in range(16):
        s = (SUBSET1_MASKS[part] >> k) & 1
        w = WEIGHTS[ibits][idx[k]]
        out[k, :3] = (ends[2 * s] * (64 - w) + ends[2 * s + 1] * w + 32) >> 6
    return out


def decode_surface(blocks, width, height):
    """The reference's output layout (block row r at byte r * width * 4) -> uint8 [height // 4 * 4, width // 4 * 4, 4]."""
    blocks = np.asarray(blocks, np.uint8).reshape(-1)
    bw, bh = width // 4, height // 4
    img = np.zeros((bh * 4, bw * 4, 4), np.uint8)
    for by in range(bh):
        for bx in range(bw):
            o = by * width * 4 + bx * 16
            img[by * 4:by * 4 + 4, bx * 4:bx * 4 + 4] = decode_block(blocks[o:o + 16]).reshape(4, 4, 4)
    return img
