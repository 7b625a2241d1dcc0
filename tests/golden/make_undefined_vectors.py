"""Hand-built deflate streams that use the codes the reference's tables do not define.

    python tests/golden/make_undefined_vectors.py      -> tests/golden/undefined_codes.json

Literal/length symbols 286 and 287 and distance codes 30, 31 (and 32..36, which a code-length
run can spill into past HDIST = 32) have no entry in /root/reference/src/const.ts:9-31.  The
reference neither rejects them nor crashes: an undefined length copies nothing (the distance
is still decoded), an undefined distance writes `len` zero bytes
(/root/reference/src/inflate.ts:98-117, 260-290).  System zlib calls all of these streams
invalid, so no stock encoder can make them; this script writes them bit by bit.

The expected output of every stream is (a) stated by hand below from reading the source and
(b) checked against tests/golden/js_model.py, a line-by-line Python model of inflate.ts with
JavaScript's `undefined` semantics.  Does not need /root/reference.
"""
import heapq
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import js_model as JS  # noqa: E402


class BitWriter:
    def __init__(self):
        self.bits = []

    def value(self, v, n):      # LSB first (header fields, extra bits)
        for k in range(n):
            self.bits.append((v >> k) & 1)

    def code(self, c, n):       # Huffman codes go MSB first
        for k in range(n - 1, -1, -1):
            self.bits.append((c >> k) & 1)

    def bytes(self):
        b = self.bits + [0] * (-len(self.bits) % 8)
        return bytes(sum(b[i + k] << k for k in range(8)) for i in range(0, len(b), 8))


def canonical(lens):
    """{symbol: length} -> {symbol: (code, length)}, by length then symbol (src/huffman.ts:8-39)"""
    out, code = {}, 0
    for l in range(min(lens.values()), max(lens.values()) + 1):
        for s in sorted(k for k, v in lens.items() if v == l):
            out[s] = (code, l)
            code += 1
        code <<= 1
    return out


def fixed_ll():
    return canonical({i: (8 if i <= 143 else 9 if i <= 255 else 7 if i <= 279 else 8) for i in range(288)})


def huffman_lengths(symbols):
    """a complete prefix code over `symbols` (equal weights) — any complete code will do for the code-length code"""
    if len(symbols) == 1:
        return {symbols[0]: 1}
    heap = [(1, i, [s]) for i, s in enumerate(symbols)]
    heapq.heapify(heap)
    depth = {s: 0 for s in symbols}
    n = len(symbols)
    while len(heap) > 1:
        a = heapq.heappop(heap)
        b = heapq.heappop(heap)
        for s in a[2] + b[2]:
            depth[s] += 1
        heapq.heappush(heap, (a[0] + b[0], n, a[2] + b[2]))
        n += 1
    return depth


def dynamic_header(w, hlit, hdist, ops):
    """ops: the code-length sequence as (symbol 0..18, extra value) pairs, exactly as they go into the stream"""
    used = sorted({s for s, _ in ops})
    cl = canonical(huffman_lengths(used))
    assert max(l for _, l in cl.values()) <= 7
    order = JS.CODELEN_VALUES
    hclen = max(i for i, s in enumerate(order) if s in cl) + 1
    hclen = max(hclen, 4)
    w.value(hlit - 257, 5)
    w.value(hdist - 1, 5)
    w.value(hclen - 4, 4)
    for i in range(hclen):
        w.value(cl[order[i]][1] if order[i] in cl else 0, 3)
    for s, ex in ops:
        w.code(*cl[s])
        if s == 16:
            w.value(ex, 2)
        elif s == 17:
            w.value(ex, 3)
        elif s == 18:
            w.value(ex, 7)


def rle_plain(lens_list):
    """code lengths -> ops using only literal lengths and 18/17 for zero runs (no symbol 16)"""
    ops, i = [], 0
    while i < len(lens_list):
        if lens_list[i] == 0:
            j = i
            while j < len(lens_list) and lens_list[j] == 0:
                j += 1
            run = j - i
            while run >= 11:
                r = min(run, 138)
                ops.append((18, r - 11))
                run -= r
            if run >= 3:
                ops.append((17, run - 3))
                run = 0
            ops += [(0, 0)] * run
            i = j
        else:
            ops.append((lens_list[i], 0))
            i += 1
    return ops


def zl(body):
    return b"\x78\x9c" + body + b"\x00\x00\x00\x00"  # the trailer is never read (src/zlib.ts:22)


def build():
    F = fixed_ll()
    vecs = []

    def fixed(name, items, expect, why):
        w = BitWriter()
        w.value(1, 1)
        w.value(1, 2)
        for it in items:
            if it[0] == "lit":
                w.code(*F[it[1]])
            elif it[0] == "eob":
                w.code(*F[256])
            else:  # ("match", length symbol, length extra value, distance code, distance extra value)
                _, ls, lex, dc, dex = it
                w.code(*F[ls])
                if ls - 257 < 29:
                    w.value(lex, JS.LENGTH_EXTRA_BIT_LEN[ls - 257])
                w.code(dc, 5)
                if dc < 30:
                    w.value(dex, JS.DISTANCE_EXTRA_BIT_LEN[dc])
        vecs.append({"name": name, "stream": zl(w.bytes()).hex(), "expect": expect.hex(), "why": why})

    fixed("fixed_len286", [("lit", 97), ("lit", 98), ("match", 286, 0, 4, 1), ("lit", 99), ("eob",)], b"abc",
          "symbol 286: length undefined -> distance code 4 and its extra bit are consumed, nothing is copied (src/inflate.ts:101-116)")
    fixed("fixed_len287", [("lit", 97), ("match", 287, 0, 0, 0), ("lit", 98), ("eob",)], b"ab",
          "symbol 287 likewise")
    fixed("fixed_dist30", [("lit", 120), ("lit", 121), ("match", 257, 0, 30, 0), ("lit", 122), ("match", 266, 1, 31, 0), ("eob",)],
          b"xy" + bytes(3) + b"z" + bytes(14),
          "distance codes 30/31: distance undefined -> repeatStartIndex is NaN, `len` zeros are written (src/inflate.ts:107-116)")
    fixed("fixed_len286_dist31", [("lit", 113), ("match", 286, 0, 31, 0), ("lit", 114), ("eob",)], b"qr",
          "both undefined: nothing read beyond the two codes, nothing written")

    # dynamic: literal/length code {a, 256, 257, 286} all 2 bits; distance code {0: 1 bit, 30: 1 bit}; HLIT = 288, HDIST = 31
    ll = {97: 2, 256: 2, 257: 2, 286: 2}
    dd = {0: 1, 30: 1}
    L, D = canonical(ll), canonical(dd)
    w = BitWriter()
    w.value(1, 1)
    w.value(2, 2)
    dynamic_header(w, 288, 31, rle_plain([ll.get(i, 0) for i in range(288)] + [dd.get(i, 0) for i in range(31)]))
    w.code(*L[97])                      # 'a'
    w.code(*L[257]); w.code(*D[0])      # len 3, dist 1 -> 'aaa'
    w.code(*L[286]); w.code(*D[0])      # undefined length: distance decoded, nothing copied
    w.code(*L[257]); w.code(*D[30])     # len 3, undefined distance -> 3 zeros
    w.code(*L[286]); w.code(*D[30])     # both undefined
    w.code(*L[97])
    w.code(*L[256])
    vecs.append({"name": "dynamic_len286_dist30", "stream": zl(w.bytes()).hex(), "expect": (b"aaaa" + bytes(3) + b"a").hex(),
                 "why": "dynamic block with lengths for symbol 286 and distance code 30 (src/inflate.ts:260-290)"})

    # dynamic, HDIST = 32 and a symbol-16 run that spills past it: distance symbols 32 and 33 get a code
    # (src/inflate.ts:187-200 pushes i++ - HLIT without looking at HDIST).  Lengths: 0 -> 1 bit; 30, 31, 32, 33 -> 3 bits.
    ll2 = {98: 1, 256: 2, 257: 2}
    ops = rle_plain([ll2.get(i, 0) for i in range(288)] + [1] + [0] * 29 + [3])  # ... distance symbols 0..30
    ops.append((16, 0))                                                           # repeat 3: symbols 31, 32, 33
    L2 = canonical(ll2)
    D2 = canonical({0: 1, 30: 3, 31: 3, 32: 3, 33: 3})
    w = BitWriter()
    w.value(1, 1)
    w.value(2, 2)
    dynamic_header(w, 288, 32, ops)
    w.code(*L2[98])                      # 'b'
    w.code(*L2[257]); w.code(*D2[33])    # len 3, distance symbol 33 (undefined) -> 3 zeros
    w.code(*L2[98])                      # 'b'
    w.code(*L2[257]); w.code(*D2[0])     # len 3, dist 1 -> 'bbb'
    w.code(*L2[257]); w.code(*D2[32])    # -> 3 zeros
    w.code(*L2[256])
    vecs.append({"name": "dynamic_dist_spill33", "stream": zl(w.bytes()).hex(), "expect": (b"b" + bytes(3) + b"bbbb" + bytes(3)).hex(),
                 "why": "a code-length run spills past HDIST = 32: distance symbols 32/33 exist in the reference's table "
                        "(src/inflate.ts:187-200) and, being undefined in const.ts, write zeros"})

    # a non-final fixed block with an undefined code, then a normal final block: decoding goes on as if nothing happened
    w = BitWriter()
    w.value(0, 1)
    w.value(1, 2)
    w.code(*F[104]); w.code(*F[105])
    w.code(*F[287]); w.code(29, 5); w.value(0x1abc & 0x1fff, 13)   # undefined length, distance code 29 with 13 extra bits consumed
    w.code(*F[256])
    w.value(1, 1)
    w.value(1, 2)
    w.code(*F[33])
    w.code(*F[256])
    vecs.append({"name": "fixed_len287_then_block", "stream": zl(w.bytes()).hex(), "expect": b"hi!".hex(),
                 "why": "the 13 extra bits of distance code 29 are consumed even though nothing is copied"})

    for v in vecs:
        got = JS.inflate(bytes.fromhex(v["stream"]))
        assert got == bytes.fromhex(v["expect"]), (v["name"], got, bytes.fromhex(v["expect"]))
    return vecs


def main():
    vecs = build()
    with open(os.path.join(HERE, "undefined_codes.json"), "w") as f:
        json.dump({"source": "hand-built; expected outputs from reading /root/reference/src/inflate.ts:98-117,260-290 and "
                             "checked with tests/golden/js_model.py", "vectors": vecs}, f, indent=1)
    print("wrote %d vectors" % len(vecs))


if __name__ == "__main__":
    main()
