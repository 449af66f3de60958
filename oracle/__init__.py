"""TEST INFRASTRUCTURE ONLY.

CPU oracle for DeSeNet's post-network inference path.  ``port`` = torch-CPU/numpy restatement,
``cport`` + ``c/oracle.c`` = plain-C restatement of the bit-exact pieces, ``ref_shim`` = in-place
import of the real reference (build container only).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this package; the
product (``desenet_b200``) never does and has no CPU fallback.
"""
