"""TEST INFRASTRUCTURE — recipe that packs the UNMODIFIED reference package into oracle/_ref/ so that the CPU arm of
bench.py (`--impl reference`, `cpu_baseline`) can time the reference's OWN modules on the GPU box, where
/root/reference does not exist.

    python oracle/build_ref.py          (also run by __graft_entry__.build() when /root/reference is present)

Output: oracle/_ref/otter_ai_ref.zip — every *.py of /root/reference/src/otter_ai (importable through zipimport,
`sys.path` entry "<zip>/src") plus the few caller files the boundary tests take functions from, byte for byte.  oracle/_ref/ is git-ignored (it never enters the history; no
reference source is copied into the repository) but it is NOT gpurun-ignored, so it travels to the GPU box like the
built .so files.  Nothing under otter_b200/ reads it; only oracle/ref_shims.py does.
"""
import hashlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OTTER_REFERENCE_SRC", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "otter_ai_ref.zip")


def build(verbose=True):
    src = os.path.join(REF, "src", "otter_ai")
    if not os.path.isdir(src):
        if verbose:
            print(f"[build_ref] {src} not found: keeping whatever oracle/_ref already holds")
        return os.path.isfile(OUT)
    files = []
    for d, _, names in os.walk(src):
        for n in sorted(names):
            if n.endswith(".py"):
                files.append(os.path.join(d, n))
    # caller-side files the boundary tests execute functions from (AST-extracted: their top-level imports need
    # accelerate / wandb / deepspeed, which this image lacks)
    for rel in ("pipeline/train/instruction_following.py", "pipeline/train/train_utils.py",
                "pipeline/demos/demo_models.py", "pipeline/mimicit_utils/mimicit_dataset.py",
                "pipeline/mimicit_utils/transforms.py"):
        f = os.path.join(REF, rel)
        if os.path.isfile(f):
            files.append(f)
    files.sort()
    digest = hashlib.sha256()
    for f in files:
        digest.update(os.path.relpath(f, REF).encode())
        digest.update(open(f, "rb").read())
    stamp = os.path.join(OUT_DIR, "otter_ai_ref.json")
    want = {"sha256": digest.hexdigest(), "files": len(files)}
    if os.path.isfile(OUT) and os.path.isfile(stamp) and json.load(open(stamp)) == want:
        return True
    os.makedirs(OUT_DIR, exist_ok=True)
    with zipfile.ZipFile(OUT, "w", zipfile.ZIP_DEFLATED) as z:
        for f in files:
            z.write(f, os.path.relpath(f, REF))          # src/otter_ai/...
    json.dump(want, open(stamp, "w"))
    if verbose:
        print(f"[build_ref] packed {len(files)} reference files into {OUT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
