"""A line-by-line Python model of the reference's inflate with JavaScript semantics.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_undefined_vectors.py to compute the
expected output of hand-built streams, and by tests/test_oracle.py as a second reading of
the source next to oracle/zlibes_oracle.c).  Small inputs only: it is bit-serial Python.

Follows /root/reference/src/inflate.ts:16-292, src/huffman.ts:8-53,
src/utils/BitReadStream.ts:1-50, src/utils/Uint8WriteStream.ts:1-25, src/zlib.ts:11-23 and
src/const.ts:9-35.  The JS behaviours that matter are modelled explicitly:

* reading an array past its end gives `undefined` (UNDEF below);
* `0 < undefined` is false, `i < undefined` is false, `x + undefined` is NaN;
* `undefined << k` is 0; storing undefined/NaN into a Uint8Array stores 0;
* `typedArray[NaN]` / `typedArray[-1]` is undefined.
"""

LENGTH_EXTRA_BIT_LEN = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
LENGTH_EXTRA_BIT_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
DISTANCE_EXTRA_BIT_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
                           8193, 12289, 16385, 24577]
DISTANCE_EXTRA_BIT_LEN = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]
CODELEN_VALUES = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]

UNDEF = None          # JavaScript `undefined`
NAN = float("nan")    # JavaScript NaN


class JsError(Exception):
    pass


class Runaway(JsError):
    """Not an error of the reference: the reference never returns on this input (see _lookup)."""


def _at(arr, i):
    """arr[i] in JS: undefined outside the array (and for a NaN index)."""
    if i != i or i is UNDEF:  # NaN
        return UNDEF
    i = int(i)
    return arr[i] if 0 <= i < len(arr) else UNDEF


def _lt(a, b):
    """a < b in JS for numbers/undefined: false when either side is undefined or NaN."""
    if a is UNDEF or b is UNDEF or a != a or b != b:
        return False
    return a < b


def _add(a, b):
    if a is UNDEF or b is UNDEF:
        return NAN
    return a + b


def _sub(a, b):
    if a is UNDEF or b is UNDEF:
        return NAN
    return a - b


class BitReadStream:  # src/utils/BitReadStream.ts
    def __init__(self, buffer, offset=0):
        self.buffer = buffer
        self.bufferIndex = offset
        v = _at(buffer, offset)
        self.nowBits = 0 if v is UNDEF else v  # undefined only matters through `& 1` / `|=`: both treat it as 0
        self.nowBitsLength = 8
        self.isEnd = False

    def read(self):  # :14-31
        if self.isEnd:
            raise JsError("Lack of data length")
        bit = self.nowBits & 1
        if self.nowBitsLength > 1:
            self.nowBitsLength -= 1
            self.nowBits >>= 1
        else:
            self.bufferIndex += 1
            if self.bufferIndex < len(self.buffer):
                self.nowBits = self.buffer[self.bufferIndex]
                self.nowBitsLength = 8
            else:
                self.nowBitsLength = 0
                self.isEnd = True
        return bit

    def readRange(self, length):  # :32-41
        while self.nowBitsLength <= length:
            self.bufferIndex += 1
            v = _at(self.buffer, self.bufferIndex)
            self.nowBits = (self.nowBits | ((0 if v is UNDEF else v) << self.nowBitsLength)) & 0xFFFFFFFF
            self.nowBitsLength += 8
        bits = self.nowBits & ((1 << length) - 1)
        self.nowBits >>= length
        self.nowBitsLength -= length
        return bits

    def readRangeCoded(self, length):  # :42-49
        bits = 0
        for _ in range(length):
            bits <<= 1
            bits |= self.read()
        return bits


class Uint8WriteStream:  # src/utils/Uint8WriteStream.ts (the growth policy is not observable)
    def __init__(self):
        self.buffer = bytearray()
        self.index = 0

    def write(self, value):
        if value is UNDEF or value != value:
            value = 0  # Uint8Array store of undefined / NaN
        self.buffer.append(int(value) & 255)
        self.index += 1


def generateHuffmanTable(codelenValues):  # src/huffman.ts:8-39
    tables = {}
    if not codelenValues:
        return tables
    codelenMin, codelenMax = min(codelenValues), max(codelenValues)
    code = 0
    for bitlen in range(codelenMin, codelenMax + 1):
        values = sorted(codelenValues.get(bitlen, []))
        table = {}
        for value in values:
            table[code] = value
            code += 1
        tables[bitlen] = table
        code <<= 1
    return tables


def makeFixedHuffmanCodelenValues():  # src/huffman.ts:41-53
    v = {7: [], 8: [], 9: []}
    for i in range(288):
        (v[8] if i <= 143 else v[9] if i <= 255 else v[7] if i <= 279 else v[8]).append(i)
    return v


FIXED_HUFFMAN_TABLE = generateHuffmanTable(makeFixedHuffmanCodelenValues())


def _lookup(stream, tables):
    """the do-while lookup of src/inflate.ts:80-93 / 157-171 / 238-252 / 267-281"""
    if not tables:
        # Math.min over no keys leaves codelenMin = Number.MAX_SAFE_INTEGER: readRangeCoded reads until read() throws
        while True:
            stream.read()
    # guard, NOT in the reference: past the end of the buffer all bits are zero and the decoder's state is its bit offset
    # inside a byte; a coded symbol that starts >= 512 bits past the end proves an endless cycle (oracle/zlibes_oracle.c)
    if not stream.isEnd and (stream.bufferIndex + 1) * 8 - stream.nowBitsLength >= len(stream.buffer) * 8 + 512:
        raise Runaway("stream never ends")
    codelenMin, codelenMax = min(tables), max(tables)
    codelen = codelenMin
    code = stream.readRangeCoded(codelenMin)
    while True:
        value = tables[codelen].get(code, UNDEF)
        if value is not UNDEF:
            return value
        if codelenMax <= codelen:
            raise JsError("Data is corrupted")
        codelen += 1
        code <<= 1
        code |= stream.read()


def _symbols(stream, buffer, dataTables, distTables):
    """src/inflate.ts:78-117 (distTables is None: fixed block) and :237-291"""
    while not stream.isEnd:
        value = _lookup(stream, dataTables)
        if value < 256:
            buffer.write(value)
            continue
        if value == 256:
            break
        repeatLengthCode = value - 257
        repeatLengthValue = _at(LENGTH_EXTRA_BIT_BASE, repeatLengthCode)
        repeatLengthExt = _at(LENGTH_EXTRA_BIT_LEN, repeatLengthCode)
        if _lt(0, repeatLengthExt):
            repeatLengthValue = _add(repeatLengthValue, stream.readRange(repeatLengthExt))
        if distTables is None:
            repeatDistanceCode = stream.readRangeCoded(5)
        else:
            repeatDistanceCode = _lookup(stream, distTables)
        repeatDistanceValue = _at(DISTANCE_EXTRA_BIT_BASE, repeatDistanceCode)
        repeatDistanceExt = _at(DISTANCE_EXTRA_BIT_LEN, repeatDistanceCode)
        if _lt(0, repeatDistanceExt):
            repeatDistanceValue = _add(repeatDistanceValue, stream.readRange(repeatDistanceExt))
        repeatStartIndex = _sub(buffer.index, repeatDistanceValue)
        i = 0
        while _lt(i, repeatLengthValue):
            buffer.write(_at(buffer.buffer, _add(repeatStartIndex, i)))
            i += 1


def _dynamic(stream, buffer):  # src/inflate.ts:120-292
    HLIT = stream.readRange(5) + 257
    HDIST = stream.readRange(5) + 1
    HCLEN = stream.readRange(4) + 4
    clv = {}
    for i in range(HCLEN):
        l = stream.readRange(3)
        if l == 0:
            continue
        clv.setdefault(l, []).append(CODELEN_VALUES[i])
    clTables = generateHuffmanTable(clv)
    dataV, distV = {}, {}
    codelen = 0
    codesNumber = HLIT + HDIST
    i = 0
    while i < codesNumber:
        rl = _lookup(stream, clTables)
        if rl == 16:
            repeat = 3 + stream.readRange(2)
        elif rl == 17:
            repeat = 3 + stream.readRange(3)
            codelen = 0
        elif rl == 18:
            repeat = 11 + stream.readRange(7)
            codelen = 0
        else:
            repeat = 1
            codelen = rl
        if codelen <= 0:
            i += repeat
        else:
            while repeat:
                if i < HLIT:
                    dataV.setdefault(codelen, []).append(i)
                else:
                    distV.setdefault(codelen, []).append(i - HLIT)
                i += 1
                repeat -= 1
    _symbols(stream, buffer, generateHuffmanTable(dataV), generateHuffmanTable(distV))


def _stored(stream, buffer):  # src/inflate.ts:42-55
    if stream.nowBitsLength < 8:
        stream.readRange(stream.nowBitsLength)
    LEN = stream.readRange(8) | stream.readRange(8) << 8
    NLEN = stream.readRange(8) | stream.readRange(8) << 8
    if LEN + NLEN != 65535:
        raise JsError("Data is corrupted")
    for _ in range(LEN):
        buffer.write(stream.readRange(8))


def inflate_raw(data: bytes, offset: int = 0) -> bytes:  # src/inflate.ts:16-40
    buffer = Uint8WriteStream()
    stream = BitReadStream(data, offset)
    bFinal = 0
    while bFinal != 1:
        bFinal = stream.readRange(1)
        bType = stream.readRange(2)
        if bType == 0:
            _stored(stream, buffer)
        elif bType == 1:
            _symbols(stream, buffer, FIXED_HUFFMAN_TABLE, None)
        elif bType == 2:
            _dynamic(stream, buffer)
        else:
            raise JsError("Not supported BTYPE : %d" % bType)
        if bFinal == 0 and stream.isEnd:
            raise JsError("Data length is insufficient")
    return bytes(buffer.buffer[:buffer.index])


def inflate(data: bytes) -> bytes:  # src/zlib.ts:11-23
    stream = BitReadStream(data)
    if stream.readRange(4) != 8:
        raise JsError("Not compressed by deflate")
    return inflate_raw(data, 2)
