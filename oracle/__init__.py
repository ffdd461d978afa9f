"""oracle/ — TEST INFRASTRUCTURE.  CPU restatement of the reference's denoising hot path (fp32 torch) and the
loader that executes the reference's own UNet files here.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product (asva_b200/, avgen/) never does.

Parity status: the reference ships no tests or golden vectors, so the oracle is pinned against the
reference ITSELF run in the authoring container (oracle/ref_loader.py imports /root/reference/avgen/models/unets
unmodified over oracle/diffusers_shim) — tests/test_host_cpu.py::test_oracle_vs_reference_live when
/root/reference is present, and the committed outputs of that run under tests/golden/ everywhere else (generator:
oracle/make_goldens.py)."""
