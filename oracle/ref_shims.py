"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference modules from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py to generate tests/golden/*.pt and by tests that pin oracle/restatement.py
against the real reference when it is present.  Nothing under otter_b200/ imports this.

Shims (harness-side only; reference files are untouched):
  1. `accelerate`, `accelerate.hooks`, `peft` stub modules (top-level imports at
     src/otter_ai/models/otter/modeling_otter.py:8,10) — inserted AFTER `import transformers`.
  2. sys.path gets /root/reference/src and /root/reference.
  3. `modeling_otter.AutoTokenizer` replaced by an offline fake (real one needs the network,
     modeling_otter.py:750,759).
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("OTTER_REFERENCE_ROOT", "/root/reference")
# the same package packed by oracle/build_ref.py (git-ignored artefact that travels to the GPU box)
REF_ZIP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "otter_ai_ref.zip")


def reference_tree_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src/otter_ai/models/otter/modeling_otter.py"))


def reference_available() -> bool:
    """The reference's python package can be imported: from /root/reference, or from oracle/_ref's archive."""
    return reference_tree_available() or os.path.isfile(REF_ZIP)


def reference_origin() -> str:
    return REF_ROOT if reference_tree_available() else (REF_ZIP if os.path.isfile(REF_ZIP) else "")


class FakeTokenizer:
    """Offline stand-in for the HF tokenizer the reference downloads in __init__."""

    def __init__(self, vocab_size=64):
        self.vocab = {}
        self.base = vocab_size
        self.pad_token = None

    @classmethod
    def from_pretrained(cls, *a, **k):
        return cls()

    def add_special_tokens(self, d):
        for tok in d.get("additional_special_tokens", []):
            self.vocab.setdefault(tok, self.base + len(self.vocab))
        if "pad_token" in d:
            self.pad_token = d["pad_token"]
            self.vocab.setdefault(d["pad_token"], self.base + len(self.vocab))
        return len(self.vocab)

    def encode(self, text):
        return [self.vocab[text]]

    def __len__(self):
        return self.base + len(self.vocab)


def _install_stubs():
    import transformers  # noqa: F401  (must come first: its find_spec probes break on stubs)

    if "accelerate" not in sys.modules:
        acc = types.ModuleType("accelerate")
        hooks = types.ModuleType("accelerate.hooks")

        class AlignDevicesHook:  # noqa: D401
            def __init__(self, *a, **k):
                pass

        def add_hook_to_module(module, hook):
            return module

        hooks.AlignDevicesHook = AlignDevicesHook
        hooks.add_hook_to_module = add_hook_to_module
        acc.hooks = hooks
        sys.modules["accelerate"] = acc
        sys.modules["accelerate.hooks"] = hooks
    if "peft" not in sys.modules:
        peft = types.ModuleType("peft")

        class LoraConfig:
            def __init__(self, *a, **k):
                pass

        class TaskType:
            CAUSAL_LM = "CAUSAL_LM"

        def get_peft_model(m, cfg):
            raise RuntimeError("peft stub: LoRA not available in the oracle harness")

        peft.LoraConfig, peft.TaskType, peft.get_peft_model = LoraConfig, TaskType, get_peft_model
        sys.modules["peft"] = peft


_CACHE = {}


def load_reference_otter():
    """Return the reference `otter_ai.models.otter.modeling_otter` module (shimmed)."""
    if "otter" in _CACHE:
        return _CACHE["otter"]
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_stubs()
    paths = (os.path.join(REF_ROOT, "src"), REF_ROOT) if reference_tree_available() else (REF_ZIP + "/src",)
    for p in paths:
        if p not in sys.path:
            sys.path.insert(0, p)
    mod = importlib.import_module("otter_ai.models.otter.modeling_otter")
    mod.AutoTokenizer = FakeTokenizer
    _CACHE["otter"] = mod
    return mod


def load_reference_flamingo():
    if "flamingo" in _CACHE:
        return _CACHE["flamingo"]
    load_reference_otter()
    mod = importlib.import_module("otter_ai.models.flamingo.modeling_flamingo")
    mod.AutoTokenizer = FakeTokenizer
    _CACHE["flamingo"] = mod
    return mod
