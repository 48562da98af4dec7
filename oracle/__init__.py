"""oracle/ — CPU restatement of the reference's algorithms.  TEST INFRASTRUCTURE ONLY: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import anything from this package.

  snark_oracle.cpp   fields, curves, Pippenger MSM, radix-2 NTT, apply-key, QAP (wasmcurves / ffjavascript), built into liboracle.so
  oracle.py          ctypes face of liboracle.so; binfile / zkey / wtns / r1cs / ptau readers; Groth16 `zkey new`, prover with
                     injected (r, s), verifier; BN254 pairing
  plonk.py           Keccak-256 transcript, PLONK prover with injected blinders, verifier, synthetic structured setup
  fflonk.py          fflonk prover with injected blinders, verifier, synthetic structured setup
  pairing_bls.py     BLS12-381 pairing (verifiers on BLS12-381 keys)
  synth_setup.py     synthetic prepared ptau + r1cs with known toxic waste (structured Groth16 keys on either curve)
  keypair.py         hash-to-G2 of the ceremony key pairs (blake2b -> ChaCha -> G2.fromRng): lets the reference's one hard-coded
                     known-answer test (test/keypar_test.js) pin the BN254 pairing and G2 arithmetic of this package

Each module's header lists the reference file:line it follows and what pins it."""
