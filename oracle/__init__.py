"""oracle/ -- CPU restatement of the reference's hot-path algorithms.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import anything from here; nothing under isopoints_b200/ does (the
product path fails loudly when libisob200.so is missing -- isopoints_b200/_ext.py).

  port.py       numpy / torch-CPU restatement of every function on the path, each citing the
                reference file:line it follows (yifita/iso-points tree).
  ref_native.py loader for the reference's OWN native extensions compiled unmodified from
                /root/reference by build_ref.py into oracle/_ref/ (frnn._C, prefix_sum, DSS._C),
                plus the host sequence of frnn.py:55-162 driving them -- the live oracle on a GPU.
  ref_python.py import hook that loads the reference's Python (DSS.models.levelset_sampling, ...)
                from /root/reference with auto-stubbed third-party packages and a declared,
                asserted 4-entry patch list for torch 2.x; used in the authoring container to
                generate tests/golden/*.npz (tests/golden/make_golden.py).  /root/reference
                does not exist on the GPU box, so nothing at test/bench time depends on it.

Pinning status (see DESIGN.md "Oracle"): port.py is checked against tests/golden/*.npz, which
were produced by the reference's own code (Python hot loop + its frnn_bf_cpu / DSS CPU twins)
run in the authoring container; on the GPU box it is additionally cross-checked against the
reference CUDA kernels in oracle/_ref.  The RGB blend (pytorch3d NormWeightedCompositor) and
pytorch3d.knn_points are third-party code absent from /root/reference: parity unpinned there.
"""
