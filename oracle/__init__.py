"""CPU oracle for the RGA3 video visual path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``rga3-release_b200/``) may import this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` use it, and there only as the
checker or as the timed CPU reference, never as the thing shipped.

Parity status: the reference repo (qirui-chen/RGA3-release) has NO tests and
NO golden vectors for this path ("parity unpinned" by the reference itself,
SURVEY.md section 8c).  The arithmetic of the path lives in third-party code
the reference imports: HuggingFace ``transformers`` (pinned 4.49.0.dev0 in the
reference's requirements.txt:25, 5.5.0 installed in this image) and Pillow.
The restatements here are therefore pinned against outputs of that code run in
this container -- see ``tests/golden/make_golden.py`` (committed generator)
and the fixtures next to it -- and, where the third-party package is part of
the image (transformers, PIL), the tests also compare against it live.

Modules
-------
index_ref     window index / cu_seqlens / rope position ids   (numpy, integer)
overlay_ref   STOM alpha-composite, warp, circle, box layers  (numpy, integer)
patchify_ref  rescale+normalise and 2x14x14 patchify          (numpy, fp32)
tower_ref     the vision tower forward in fp32                (torch, fp32)
hf_ref        builders for the real HF tower / video processor with the
              seeded weights the tests and the bench use
"""
