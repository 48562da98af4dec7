// snarkb200.mjs — overrides the bulk methods of snarkjs' curve singleton with the B200 backend.
// Usage:  import { useB200 } from "./snarkb200.mjs";  await useB200("bn128");  await snarkjs.groth16.prove(zkey, wtns);
// Not runnable in this repository's image (no Node.js); see INTEGRATION.md for the line-by-line rationale.
import { createRequire } from "module";
import * as snarkjs from "snarkjs";
const addon = createRequire(import.meta.url)("./build/Release/snarkb200_napi.node");

const flatten = (b) => (b instanceof Uint8Array ? b : Buffer.concat(b.buffers));          // BigBuffer (ffjavascript 12692-12778)

export async function useB200(curveName = "bn128", device = 0) {
  const curve = await snarkjs.curves.getCurveFromName(curveName);                           // the singleton snarkjs itself gets (src/curves.js:36-53)
  const ctx = addon.createContext(curveName === "bn128" ? 0 : 1, device);
  const n8q = curve.F1.n8;
  for (const [G, gid] of [[curve.G1, 1], [curve.G2, 2]]) {
    G.multiExpAffine = async (buffBases, buffScalars) => {
      const bases = flatten(buffBases), scalars = flatten(buffScalars);
      const n = Math.floor(bases.byteLength / (G.F.n8 * 2));
      if (n == 0) return G.zero;
      if (Math.floor(scalars.byteLength / n) * n != scalars.byteLength) throw new Error("Scalar size does not match");
      return new Uint8Array(await addon.multiExpAffine(ctx, gid, bases, scalars, n8q));
    };
  }
  const Fr = curve.Fr;
  const fft = (inverse) => async (buff) => {
    const isArray = Array.isArray(buff);
    const b = isArray ? Buffer.concat(buff) : flatten(buff);
    const n = b.byteLength / Fr.n8;
    if (!Number.isInteger(Math.log2(n))) throw new Error("fft must be multiple of 2");
    const out = new Uint8Array(await addon.nttFr(ctx, b, inverse));
    return isArray ? Array.from({ length: n }, (_, i) => out.slice(i * Fr.n8, (i + 1) * Fr.n8)) : out;
  };
  Fr.fft = fft(0);
  Fr.ifft = fft(1);
  Fr.batchApplyKey = async (buff, first, inc) => new Uint8Array(await addon.frBatchApplyKey(ctx, flatten(buff), Fr.e(first), Fr.e(inc)));
  Fr.batchToMontgomery = async (buff) => new Uint8Array(await addon.frConvert(ctx, flatten(buff), 1));
  Fr.batchFromMontgomery = async (buff) => new Uint8Array(await addon.frConvert(ctx, flatten(buff), 0));
  const queueAction = curve.tm.queueAction.bind(curve.tm);
  curve.tm.queueAction = async (task) => {                                                 // joinABC's raw tasks (src/groth16_prove.js:338-355)
    const call = task.find((t) => t.cmd == "CALL");
    if (call && call.fnName == "qap_joinABC") {
      const [a, b, c] = task.filter((t) => t.cmd == "ALLOCSET").map((t) => t.buff);
      return [new Uint8Array(await addon.qapJoinAbc(ctx, a, b, c))];
    }
    return queueAction(task);
  };
  return { curve, ctx, addon };
}
