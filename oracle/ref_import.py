"""TEST INFRASTRUCTURE ONLY. Makes the UNMODIFIED reference at /root/reference importable in the build container
(SURVEY.md section 8c): stub modules for ipdb / ftfy / kornia on sys.path ahead of the reference, and the four
transformers-4.x names model/iqm.py:12-17,644 needs. Used by oracle/gen_golden.py to pin the restatement in
oracle/aaclip_oracle.py. /root/reference does not exist on the GPU box: nothing executed there imports this."""
import os
import sys

REF_ROOT = os.environ.get("AACLIP_REFERENCE", "/root/reference")
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def install():
    """Put the shims and the reference on sys.path and patch transformers; returns the reference root."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    for p in (REF_ROOT, SHIM):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, SHIM)
    # the repository root carries shim modules under the reference's import names (model/, dataset/, forward_utils.py,
    # utils.py): make sure `import model` below resolves to the REFERENCE, also if a shim was imported earlier
    for name in list(sys.modules):
        if name.split(".")[0] in ("model", "dataset", "forward_utils", "utils"):
            f = getattr(sys.modules[name], "__file__", "") or ""
            if not f.startswith(REF_ROOT):
                del sys.modules[name]
    # the reference's model/ directory has no __init__.py (a namespace package), and a REGULAR package of the same name
    # later on sys.path (our shim) would win over it whatever the path order: bind the name to the reference explicitly
    if "model" not in sys.modules:
        import types
        pkg = types.ModuleType("model")
        pkg.__path__ = [os.path.join(REF_ROOT, "model")]
        sys.modules["model"] = pkg
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu

    if not hasattr(mu, "apply_chunking_to_forward"):
        mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    if not hasattr(mu, "prune_linear_layer"):
        mu.prune_linear_layer = pu.prune_linear_layer
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        def _no_pruning(*a, **k):
            raise NotImplementedError("head pruning is not used on the AA-CLIP path")
        mu.find_pruneable_heads_and_indices = _no_pruning
    if not hasattr(mu.PreTrainedModel, "get_head_mask"):
        mu.PreTrainedModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n
    return REF_ROOT
