"""Regenerates tests/golden/ from the reference's own test vectors.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):  python tests/golden/make_golden.py

Outputs
  ref_fixture_compressed.zlib  — byte copy of /root/reference/test/data/compressed.bin
                                 (test DATA, not source; inflating it with system
                                 zlib reproduces test/data/raw.bin, whose sha256 is
                                 recorded in vectors.json)
  vectors.json                 — the inline known-answer vectors of
                                 /root/reference/test/index.js:7-10,89-94, the
                                 fixture digests, and the SURVEY.md §8c cross-check
                                 table (independent model of the reference).
"""
import hashlib
import json
import os
import shutil
import zlib

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

RAW = [84, 104, 105, 115, 32, 105, 115, 32, 122, 108, 105, 98, 46, 101, 115]
UNCOMPRESSED = [120, 156, 1, 15, 0, 240, 255, 84, 104, 105, 115, 32, 105, 115, 32, 122, 108, 105, 98, 46, 101, 115, 43, 35, 5, 108]
FIXED = [120, 156, 11, 201, 200, 44, 86, 0, 162, 170, 156, 204, 36, 189, 212, 98, 0, 43, 35, 5, 108]
DYNAMIC = [120, 156, 13, 194, 65, 9, 0, 0, 8, 3, 192, 42, 38, 48, 141, 9, 4, 193, 129, 191, 253, 150, 126, 194, 213, 130, 241, 116, 232, 28, 26, 43, 35, 5, 108]
# test/index.js:89 — the JS literal drops the backslashes of \' and \], so the
# alphabet has 93 characters (no backslash); repeated until >= 1000 chars.
ASCII = '!"#$%&\'()*+,-./0123456789:;<=>?@ABCDEFGHIJKLMNOPQRSTUVWXYZ[]^_`abcdefghijklmnopqrstuvwxyz{|}~'

# generator, n, adler32, deflate size, sha256(out)[:16], tokens per block
MODEL_TABLE = [
    ["RAW", 15, "2b23056c", 35, "da00ac58cbb8a51f", [13]],
    ["REPEAT", 1023, "0b8d3d37", 245, "90cc9c96c2585ccc", [252]],
    ["FIXTURE", 480400, "140fa15b", 191734, "45b216a3f0348dc7", [18082, 28161, 33565, 23037]],
    ["G1", 2, "0012000e", 21, "ef1d0bcf45fd5882", [2]],
    ["G1", 3, "0031001f", 21, "e6a5c8288823d7fb", [3]],
    ["G1", 4096, "9a15f86a", 559, "a877a171e1eea69a", [498]],
    ["G1", 131072, "94e30ef2", 854, "19615bc27cc97883", [773]],
    ["G1", 131074, "b2d70eff", 868, "05f31d427579b5ed", [773, 2]],
    ["G1", 300000, "5f5bc63a", 2389, "0df102c1b6028aea", [773, 773, 591]],
    ["G2", 4096, "10000001", 53, "b7f3c499d2c040af", [241]],
    ["G2", 131072, "001e0001", 149, "7c02b0e2ecc5a81e", [516]],
    ["G3", 4096, "742bf93d", 4153, "6f632c5676384a34", [4094]],
    ["G3", 65536, "04f60207", 65636, "12af5d9acf69c7bc", [65226]],
    ["G4", 4096, "f4a0205a", 87, "4aaa999614acf286", [241]],
    ["G5", 4096, "2e5c88ee", 2347, "533c80a92b95a0ad", [2472]],
    ["G5", 200000, "b8f1f649", 117570, "95c155561c4d1a88", [40557, 22293]],
]


def main():
    raw = open(os.path.join(REF, "test/data/raw.bin"), "rb").read()
    cmp_ = open(os.path.join(REF, "test/data/compressed.bin"), "rb").read()
    assert zlib.decompress(cmp_) == raw
    shutil.copyfile(os.path.join(REF, "test/data/compressed.bin"), os.path.join(HERE, "ref_fixture_compressed.zlib"))
    rep = ""
    while len(rep) < 1000:
        rep += ASCII
    doc = {
        "source": "zprodev/zlib.es v0.6.0 test/index.js + test/data",
        "RAW": RAW, "UNCOMPRESSED": UNCOMPRESSED, "FIXED": FIXED, "DYNAMIC": DYNAMIC,
        "REPEAT_ALPHABET": ASCII, "REPEAT_LEN": len(rep),
        "fixture": {
            "raw_len": len(raw), "raw_sha256": hashlib.sha256(raw).hexdigest(),
            "raw_adler32": "%08x" % zlib.adler32(raw),
            "compressed_len": len(cmp_), "compressed_sha256": hashlib.sha256(cmp_).hexdigest(),
        },
        "adler_kat": {"RAW": "2b23056c", "FIXTURE": "140fa15b"},
        "throwing_lengths": [0, 1, 131073],
        "model_table": MODEL_TABLE,
    }
    with open(os.path.join(HERE, "vectors.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", HERE)


if __name__ == "__main__":
    main()
