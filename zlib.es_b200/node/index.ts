// index.ts — TypeScript host side of the drop-in: same two exports, same signatures as
// zlib.es (/root/reference/dist/tsc/zlib.d.ts:4-5, /root/reference/src/zlib.ts:11,25).
// All work happens in the N-API addon (addon.c -> libzles.so -> sm_100a kernels); there is no
// JavaScript fallback: if the addon cannot be loaded, importing this module throws.
//
//   import { inflate, deflate } from 'zles-b200';      // was: from 'zlib.es'
//
// eslint-disable-next-line @typescript-eslint/no-var-requires
const native = require('./zles.node') as {
  deflate(input: Uint8Array): Uint8Array;
  inflate(input: Uint8Array): Uint8Array;
  deflateBatch(inputs: Uint8Array[]): Uint8Array[];
  inflateBatch(inputs: Uint8Array[]): Uint8Array[];
  deflateRaw(input: Uint8Array): Uint8Array;
  inflateRaw(input: Uint8Array): Uint8Array;
  gzip(input: Uint8Array): Uint8Array;
  gunzip(input: Uint8Array): Uint8Array;
};

/** zlib-wrapped DEFLATE of `input` (CMF/FLG 78 9C, Adler-32 trailer), computed on the GPU. */
export function deflate(input: Uint8Array): Uint8Array {
  return native.deflate(input);
}

/** Inverse of deflate; accepts any zlib stream the reference accepts and throws the same Error messages. */
export function inflate(input: Uint8Array): Uint8Array {
  return native.inflate(input);
}

/** `inputs.map(deflate)` in one launch sequence (small-message / PNG-row shape). */
export function deflateBatch(inputs: Uint8Array[]): Uint8Array[] {
  return native.deflateBatch(inputs);
}

/** `inputs.map(inflate)` in one launch sequence. */
export function inflateBatch(inputs: Uint8Array[]): Uint8Array[] {
  return native.inflateBatch(inputs);
}

/** Raw DEFLATE data (RFC 1951), no container: what the reference's deflate core returns (src/deflate.ts:14). */
export function deflateRaw(input: Uint8Array): Uint8Array {
  return native.deflateRaw(input);
}

/** Inverse of deflateRaw: the reference's inflate core with offset 0 (src/inflate.ts:16). */
export function inflateRaw(input: Uint8Array): Uint8Array {
  return native.inflateRaw(input);
}

/** gzip container (RFC 1952: CRC-32 and length trailer) around the same DEFLATE data. */
export function gzip(input: Uint8Array): Uint8Array {
  return native.gzip(input);
}

/** Inverse of gzip; verifies CRC-32 and length. */
export function gunzip(input: Uint8Array): Uint8Array {
  return native.gunzip(input);
}
