"""Synthetic .CRN file writer (TEST / BENCH INFRASTRUCTURE).

Builds valid CRN streams of any size the 16-bit header allows (the reference *compressor* refuses levels
above 4096 px, SURVEY D6, but its decoder does not), with random palettes and block indices, so the
transcoder can be exercised at BASELINE.json's 8192x8192 DXT5 size and compared bit-for-bit with the
reference decoder (oracle/_ref) and the oracle port.  Wire format as documented in SURVEY.md Appendix A
(reference inc/crn_defs.h:286-341, inc/crn_decomp.h:3044-3123, :3715-3851, :3944-4223; writer side
crnlib/crn_comp.cpp:43-123, :231-261, :295-422, :1356-1496).  The CRC fields are filled with the
reference's crc16 (crnlib/crn_checksum.cpp) so crnd_validate_file accepts the result."""
import heapq

import numpy as np

FMT = dict(DXT1=0, DXT5=2, DXN_XY=7, DXN_YX=8, DXT5A=9)
ORDER = [17, 18, 19, 20, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15, 16]
DXT5_FROM_LINEAR = [0, 2, 3, 4, 5, 6, 7, 1]


def huff_lengths(freq, max_len=16):
    """Code lengths of a Huffman code limited to max_len bits (frequencies halved until it fits)."""
    freq = np.asarray(freq, np.int64).copy()
    n = len(freq)
    used = np.nonzero(freq)[0]
    lens = np.zeros(n, np.uint8)
    if len(used) == 0:
        return lens
    if len(used) == 1:
        lens[used[0]] = 1
        return lens
    while True:
        heap = [(int(freq[i]), int(i), None, None) for i in used]
        heapq.heapify(heap)
        nxt = n
        nodes = {}
        while len(heap) > 1:
            a = heapq.heappop(heap); b = heapq.heappop(heap)
            nodes[nxt] = (a[1], b[1])
            heapq.heappush(heap, (a[0] + b[0], nxt, None, None))
            nxt += 1
        depth = {heap[0][1]: 0}
        stack = [heap[0][1]]
        out = np.zeros(n, np.int64)
        while stack:
            k = stack.pop()
            if k in nodes:
                for c in nodes[k]:
                    depth[c] = depth[k] + 1
                    stack.append(c)
            else:
                out[k] = depth[k]
        if out.max() <= max_len:
            return out.astype(np.uint8)
        freq[used] = (freq[used] + 1) // 2


def canonical_codes(lens):
    """Canonical code values: shorter first, ties by symbol index (inc/crn_decomp.h:2161-2235)."""
    lens = np.asarray(lens, np.int64)
    codes = np.zeros(len(lens), np.int64)
    code = 0
    for l in range(1, 17):
        idx = np.nonzero(lens == l)[0]
        codes[idx] = code + np.arange(len(idx))
        code = (code + len(idx)) << 1
    return codes


class BitWriter:
    def __init__(self):
        self.codes = []
        self.lens = []

    def put(self, value, nbits):
        self.codes.append(np.asarray([value], np.int64))
        self.lens.append(np.asarray([nbits], np.int64))

    def put_array(self, values, nbits):
        self.codes.append(np.asarray(values, np.int64).ravel())
        self.lens.append(np.asarray(nbits, np.int64).ravel())

    def bytes(self):
        codes = np.concatenate(self.codes) if self.codes else np.zeros(0, np.int64)
        lens = np.concatenate(self.lens) if self.lens else np.zeros(0, np.int64)
        keep = lens > 0
        codes, lens = codes[keep], lens[keep]
        total = int(lens.sum())
        nbytes = (total + 7) // 8 + 1        # the reference pads 7 zero bits, then flushes whole bytes
        out = np.zeros(nbytes + 4, np.uint8)
        if total:
            start = np.concatenate([[0], np.cumsum(lens)[:-1]])
            # value left-aligned inside a 32-bit window starting at the byte containing `start`
            sh = 32 - (start & 7) - lens
            v = (codes << sh).astype(np.uint64)
            b0 = (start >> 3).astype(np.int64)
            # codes never overlap, so OR == sum: accumulate the four byte lanes with bincount (exact in float64)
            acc = np.zeros(nbytes + 4, np.float64)
            for k in range(4):
                acc += np.bincount(b0 + k, weights=((v >> np.uint64(24 - 8 * k)) & np.uint64(0xFF)).astype(np.float64), minlength=nbytes + 4)
            out = acc.astype(np.uint8)
        return out[:nbytes].tobytes()


def send_model(bw, lens):
    """Transmit a static Huffman model (decode_receive_static_data_model in reverse)."""
    lens = np.asarray(lens, np.int64)
    used = np.nonzero(lens)[0]
    total = int(used[-1]) + 1 if len(used) else 0
    bw.put(total, 14)
    if not total:
        return
    # run-length tokens over lens[:total]
    toks = []   # (symbol, extra value, extra bits)
    i = 0
    while i < total:
        v = int(lens[i]); j = i
        while j < total and lens[j] == v:
            j += 1
        run = j - i
        if v == 0:
            while run >= 11:
                r = min(run, 138); toks.append((18, r - 11, 7)); run -= r
            if run >= 3:
                toks.append((17, run - 3, 3)); run = 0
            toks += [(0, 0, 0)] * run
        else:
            toks.append((v, 0, 0)); run -= 1
            while run >= 7:
                r = min(run, 70); toks.append((20, r - 7, 6)); run -= r
            if run >= 3:
                toks.append((19, run - 3, 2)); run = 0
            toks += [(v, 0, 0)] * run
        i = j
    f = np.zeros(21, np.int64)
    for s, _, _ in toks:
        f[s] += 1
    cl = huff_lengths(f, 7)
    ncl = max(k + 1 for k in range(21) if cl[ORDER[k]])
    bw.put(ncl, 5)
    for k in range(ncl):
        bw.put(int(cl[ORDER[k]]), 3)
    cc = canonical_codes(cl)
    for s, ev, eb in toks:
        bw.put(int(cc[s]), int(cl[s]))
        if eb:
            bw.put(ev, eb)


class Model:
    def __init__(self, symbols, nsyms):
        f = np.bincount(np.asarray(symbols, np.int64).ravel(), minlength=nsyms)
        self.lens = huff_lengths(f)
        self.codes = canonical_codes(self.lens)

    def encode(self, symbols):
        s = np.asarray(symbols, np.int64)
        return self.codes[s], self.lens[s].astype(np.int64)


def _crc16_ref(data):
    # const uint16 q = *pBuf++ ^ (crc >> 8); crc <<= 8; uint16 r = (q >> 4) ^ q; crc ^= r; r <<= 5; crc ^= r; r <<= 7; crc ^= r;
    crc = 0xFFFF   # ~0
    for b in data:
        q = (b ^ (crc >> 8)) & 0xFFFF
        crc = (crc << 8) & 0xFFFF
        r = ((q >> 4) ^ q) & 0xFFFF
        crc ^= r
        r = (r << 5) & 0xFFFF
        crc ^= r
        r = (r << 7) & 0xFFFF
        crc ^= r
    return (~crc) & 0xFFFF


def _skewed(rng, n, size, p=None):
    """Indices with a long-tailed distribution so the Huffman codes have a spread of lengths.  p is the
    geometric parameter: the default 8/n gives a flat, high-entropy stream (many codes longer than the
    decoder's first-level table), p ~ 0.05 gives bitrates like real .crn files (1-1.3 bits/texel)."""
    x = rng.geometric(min(0.5, 8.0 / max(n, 8)) if p is None else p, size) - 1
    perm = rng.permutation(n)
    return perm[x % n]


def synth_crn(width, height, fmt="DXT5", levels=None, faces=1, seed=1, n_color_ep=1024, n_color_sel=1024, n_alpha_ep=512, n_alpha_sel=512,
              with_crc=True, skew=None):
    rng = np.random.default_rng(seed)
    f = FMT[fmt]
    has_color = f in (0, 2)
    has_a0 = f != 0
    has_a1 = f in (7, 8)
    if levels is None:
        levels = 1
        while (max(width, height) >> levels) > 0:
            levels += 1
    # ---- palettes
    segs = {}
    if has_color:
        ce = rng.integers(0, [32, 64, 32, 32, 64, 32], (n_color_ep, 6))
        bw = BitWriter()
        d = np.diff(np.vstack([np.zeros((1, 6), np.int64), ce]), axis=0)
        d5 = (d[:, [0, 2, 3, 5]] % 32); d6 = (d[:, [1, 4]] % 64)
        m0, m1 = Model(d5, 32), Model(d6, 64)
        send_model(bw, m0.lens); send_model(bw, m1.lens)
        syms = np.empty((n_color_ep, 6), np.int64); lens = np.empty((n_color_ep, 6), np.int64)
        for col, (m, src, k) in enumerate([(m0, d5, 0), (m1, d6, 0), (m0, d5, 1), (m0, d5, 2), (m1, d6, 1), (m0, d5, 3)]):
            syms[:, col], lens[:, col] = m.encode(src[:, k])
        bw.put_array(syms, lens)
        segs["ce"] = bw.bytes()
        # selectors: 8 nibbles per entry, XOR-delta coded on the 32-bit linear value
        lin = rng.integers(0, 1 << 32, n_color_sel, dtype=np.uint64).astype(np.int64)
        prev = np.concatenate([[0], lin[:-1]])
        x = lin ^ prev
        nib = np.stack([(x >> (4 * j)) & 15 for j in range(8)], axis=1)
        ms = Model(nib, 16)
        bw = BitWriter(); send_model(bw, ms.lens)
        c, l = ms.encode(nib); bw.put_array(c, l)
        segs["cs"] = bw.bytes()
    if has_a0:
        ae = rng.integers(0, 256, (n_alpha_ep, 2))
        d = np.diff(np.vstack([np.zeros((1, 2), np.int64), ae]), axis=0) % 256
        m = Model(d, 256)
        bw = BitWriter(); send_model(bw, m.lens)
        c, l = m.encode(d); bw.put_array(c, l)
        segs["ae"] = bw.bytes()
        lin = rng.integers(0, 64, (n_alpha_sel, 8))
        x = lin ^ np.vstack([np.zeros((1, 8), np.int64), lin[:-1]])
        m = Model(x, 64)
        bw = BitWriter(); send_model(bw, m.lens)
        c, l = m.encode(x); bw.put_array(c, l)
        segs["as"] = bw.bytes()
    # ---- per-level block symbols
    level_syms = []
    all_ref, all_ced, all_aed, all_cs, all_as = [], [], [], [], []
    for lv in range(levels):
        w, h = max(1, width >> lv), max(1, height >> lv)
        bx, by = (w + 3) // 4, (h + 3) // 4
        W, H = (bx + 1) & ~1, (by + 1) & ~1
        n = faces * H * W
        ref = rng.choice(3, size=(faces * H, W), p=[0.5, 0.3, 0.2])
        rows = faces * H
        # endpoint indices: row by row so that "left" (1) and "top" (2) references are consistent
        def resolve(npal):
            """Endpoint index of every block.  New-endpoint blocks (ref 0) move the running index by a skewed
            delta, "left" (1) keeps it, "top" (2) reloads the index of the block above: a segmented prefix
            sum per row, carried across rows and faces exactly like the decoder's running index."""
            e = np.zeros((rows, W), np.int64)
            step = _skewed(rng, npal, rows * W, skew).reshape(rows, W)
            carry = 0
            ar = np.arange(W)
            for r in range(rows):
                rr = ref[r]
                if r % H == 0 and (rr == 2).any():
                    rr = rr.copy(); rr[rr == 2] = 0          # no row above inside this face for the writer to reference
                    ref[r] = rr
                c = np.cumsum(np.where(rr == 0, step[r], 0))
                h = np.maximum.accumulate(np.where(rr == 2, ar, -1))
                top = e[r - 1] if r else e[0]
                hh = np.maximum(h, 0)
                row = np.where(h >= 0, top[hh] + c - c[hh], carry + c) % npal
                e[r] = row
                carry = int(row[-1])
            return e
        ce_idx = resolve(n_color_ep) if has_color else None
        a0_idx = resolve(n_alpha_ep) if has_a0 else None
        a1_idx = resolve(n_alpha_ep) if has_a1 else None
        # NOTE: resolve() may rewrite ref on the first row of a face; all components share `ref`, so run once more
        # with the final ref for consistency of the earlier components
        if has_color:
            ce_idx = resolve(n_color_ep)
        if has_a0:
            a0_idx = resolve(n_alpha_ep)
        if has_a1:
            a1_idx = resolve(n_alpha_ep)

        def deltas(e, npal):
            flat = e.ravel()
            prev = np.concatenate([[0], flat[:-1]])
            return (flat - prev) % npal
        lvl = dict(W=W, H=H, n=n, ref=ref.ravel())
        if has_color:
            lvl["ced"] = deltas(ce_idx, n_color_ep); lvl["cs"] = _skewed(rng, n_color_sel, n, skew)
            all_ced.append(lvl["ced"][lvl["ref"] == 0]); all_cs.append(lvl["cs"])
        if has_a0:
            lvl["a0d"] = deltas(a0_idx, n_alpha_ep); lvl["s0"] = _skewed(rng, n_alpha_sel, n, skew)
            all_aed.append(lvl["a0d"][lvl["ref"] == 0]); all_as.append(lvl["s0"])
        if has_a1:
            lvl["a1d"] = deltas(a1_idx, n_alpha_ep); lvl["s1"] = _skewed(rng, n_alpha_sel, n, skew)
            all_aed.append(lvl["a1d"][lvl["ref"] == 0]); all_as.append(lvl["s1"])
        r2 = ref.reshape(faces, H, W)
        grp = (r2[:, 0::2, 0::2] | (r2[:, 1::2, 0::2] << 2) | (r2[:, 0::2, 1::2] << 4) | (r2[:, 1::2, 1::2] << 6))
        lvl["grp"] = grp
        all_ref.append(grp.ravel())
        level_syms.append(lvl)
    m_ref = Model(np.concatenate(all_ref), 256)
    m_ce = Model(np.concatenate(all_ced), n_color_ep) if has_color else None
    m_cs = Model(np.concatenate(all_cs), n_color_sel) if has_color else None
    m_ae = Model(np.concatenate(all_aed), n_alpha_ep) if has_a0 else None
    m_as = Model(np.concatenate(all_as), n_alpha_sel) if has_a0 else None
    bw = BitWriter()
    send_model(bw, m_ref.lens)
    if has_color:
        send_model(bw, m_ce.lens); send_model(bw, m_cs.lens)
    if has_a0:
        send_model(bw, m_ae.lens); send_model(bw, m_as.lens)
    segs["tables"] = bw.bytes()
    # ---- level bitstreams: per block up to 7 symbol slots in stream order
    level_bytes = []
    for lvl in level_syms:
        W, H, n = lvl["W"], lvl["H"], lvl["n"]
        slots = []   # (codes, lens) arrays of shape (n,)
        ref = lvl["ref"]
        yy = (np.arange(n) // W) % H
        xx = np.arange(n) % W
        at_group = ((yy & 1) == 0) & ((xx & 1) == 0)
        gc = np.zeros(n, np.int64); gl = np.zeros(n, np.int64)
        c, l = m_ref.encode(lvl["grp"].ravel())
        gc[at_group] = c; gl[at_group] = l
        slots.append((gc, gl))
        z = ref == 0

        def cond(model, syms):
            c, l = model.encode(syms)
            return np.where(z, c, 0), np.where(z, l, 0)
        if has_color:
            slots.append(cond(m_ce, lvl["ced"]))
        if has_a0:
            slots.append(cond(m_ae, lvl["a0d"]))
        if has_a1:
            slots.append(cond(m_ae, lvl["a1d"]))
        if has_color:
            slots.append(m_cs.encode(lvl["cs"]))
        if has_a0:
            slots.append(m_as.encode(lvl["s0"]))
        if has_a1:
            slots.append(m_as.encode(lvl["s1"]))
        codes = np.stack([s[0] for s in slots], axis=1)
        lens = np.stack([s[1] for s in slots], axis=1)
        bw = BitWriter(); bw.put_array(codes, lens)
        level_bytes.append(bw.bytes())
    # ---- assemble
    header_size = 70 + 4 * levels
    body = bytearray()
    ofs = header_size

    def place(seg):
        nonlocal ofs
        o = ofs
        body.extend(seg); ofs += len(seg)
        return o
    pal = []
    for key, num in (("ce", n_color_ep), ("cs", n_color_sel), ("ae", n_alpha_ep), ("as", n_alpha_sel)):
        if key in segs:
            pal.append((place(segs[key]), len(segs[key]), num))
        else:
            pal.append((0, 0, 0))
    tables_ofs = place(segs["tables"]); tables_size = len(segs["tables"])
    lofs = [place(b) for b in level_bytes]
    data_size = ofs

    def be(v, n):
        return int(v).to_bytes(n, "big")
    hdr_tail = be(data_size, 4) + b"\0\0" + be(width, 2) + be(height, 2) + be(levels, 1) + be(faces, 1) + be(f, 1) + be(0, 2) + be(0, 4) + be(0, 4) + be(0, 4)
    for (o, s, nn) in pal:
        hdr_tail += be(o, 3) + be(s, 3) + be(nn, 2)
    hdr_tail += be(tables_size, 2) + be(tables_ofs, 3)
    for o in lofs:
        hdr_tail += be(o, 4)
    data_crc = _crc16_ref(bytes(body)) if with_crc else 0
    hdr_tail = hdr_tail[:4] + be(data_crc, 2) + hdr_tail[6:]
    header_crc = _crc16_ref(hdr_tail)
    header = be((ord("H") << 8) | ord("x"), 2) + be(header_size, 2) + be(header_crc, 2) + hdr_tail
    assert len(header) == header_size
    return bytes(header) + bytes(body)
