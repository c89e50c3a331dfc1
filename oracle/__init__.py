"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's hot-path arithmetic (DVIS_Plus @ c0eb2495), used as the checker by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product
package `dvis_plus_b200` never imports anything from here; it fails loudly when its CUDA library is
missing instead of falling back to this code.

  oracle.c_oracle   -- ctypes wrappers over oracle/msda_oracle.c (plain C; MSDA fwd/bwd, mask logits)
  oracle.torch_port -- torch-CPU restatement of the module-level path (MSDeformAttn, encoder, pixel
                       decoder, mask head, tracker, refiner); same algorithm as the reference's own CPU
                       path (F.grid_sample / einsum / multi-head attention), timed as `cpu_baseline`.

Parity pin: tests/test_oracle.py checks both against golden vectors produced by importing the
unmodified reference in the build container (tests/golden/make_golden.py).
"""
