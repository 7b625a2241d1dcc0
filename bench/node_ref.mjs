// bench/node_ref.mjs — times the UNMODIFIED zlib.es (the reference) on the host CPU with Node: single-threaded and with
// one worker_thread per core over independent 128 KiB-aligned pieces (BASELINE.json north_star, SURVEY.md 8d item 4).
// Never run in this repo's CI: neither the build image nor the GPU boxes have Node.  bench.py --impl reference calls it
// only when `node` is on PATH and ZLIBES_DIST points at the reference's dist/cjs/zlib.js; otherwise it times the C
// restatement of the same algorithm (oracle/).
//
// usage: node bench/node_ref.mjs <path to zlib.es dist/cjs/zlib.js> <input file> [workers = cores]
// prints one JSON line: {"single": {deflate_gbs, inflate_gbs}, "workers": {n, deflate_gbs, inflate_gbs}, "ratio": ...}
import { Worker, isMainThread, parentPort, workerData } from 'node:worker_threads';
import { readFileSync } from 'node:fs';
import { createRequire } from 'node:module';
import { cpus } from 'node:os';

const require = createRequire(import.meta.url);
const CHUNK = 131072;

function pieces(n, k) {
  const nchunks = Math.ceil(n / CHUNK);
  const out = [];
  for (let w = 0; w < k; w++) {
    const a = Math.min(n, Math.floor((nchunks * w) / k) * CHUNK);
    const b = w === k - 1 ? n : Math.min(n, Math.floor((nchunks * (w + 1)) / k) * CHUNK);
    if (b > a) out.push([a, b]);
  }
  return out;
}

if (!isMainThread) {
  const { lib, file, a, b } = workerData;
  const zl = require(lib);
  const data = new Uint8Array(readFileSync(file)).subarray(a, b);
  let t0 = performance.now();
  const z = zl.deflate(data);
  const td = performance.now() - t0;
  t0 = performance.now();
  const back = zl.inflate(z);
  const ti = performance.now() - t0;
  parentPort.postMessage({ td, ti, comp: z.length, ok: back.length === data.length });
} else {
  const [lib, file, wArg] = process.argv.slice(2);
  if (!lib || !file) {
    console.error('usage: node bench/node_ref.mjs <zlib.es dist/cjs/zlib.js> <input file> [workers]');
    process.exit(2);
  }
  const zl = require(lib);
  const data = new Uint8Array(readFileSync(file));
  const n = data.length;
  // single-threaded, on a bounded sample (the reference deflates ~10 MB/s)
  const sample = data.subarray(0, Math.min(n, 8 * CHUNK * 8));
  let t0 = performance.now();
  const z = zl.deflate(sample);
  const td = (performance.now() - t0) / 1e3;
  t0 = performance.now();
  const back = zl.inflate(z);
  const ti = (performance.now() - t0) / 1e3;
  const single = { bytes: sample.length, deflate_gbs: sample.length / td / 1e9, inflate_gbs: sample.length / ti / 1e9, ok: back.length === sample.length };
  // one worker per core, each on its own contiguous run of whole chunks
  const k = Number(wArg) || cpus().length;
  const parts = pieces(n, k);
  const wall0 = performance.now();
  const results = await Promise.all(parts.map(([a, b]) => new Promise((resolve, reject) => {
    const w = new Worker(new URL(import.meta.url), { workerData: { lib, file, a, b } });
    w.once('message', resolve);
    w.once('error', reject);
  })));
  const wall = (performance.now() - wall0) / 1e3;
  const tdMax = Math.max(...results.map((r) => r.td)) / 1e3;
  const tiMax = Math.max(...results.map((r) => r.ti)) / 1e3;
  const comp = results.reduce((s, r) => s + r.comp, 0);
  console.log(JSON.stringify({
    single,
    workers: { n: parts.length, deflate_gbs: n / tdMax / 1e9, inflate_gbs: n / tiMax / 1e9, wall_s: wall, ok: results.every((r) => r.ok) },
    ratio: n / comp,
  }));
}
