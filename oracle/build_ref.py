"""Recipe for oracle/_ref/: the reference's OWN hot-path modules for bench.py's CPU arm (`cpu_baseline.kind = "reference"`).

The reference has no compiled code; its path is three pure-Python files (plus the training script that drives them).  They
are copied VERBATIM from /root/reference
into the git-ignored oracle/_ref/ (never into the repository history) so that they travel to the GPU box with the snapshot,
where /root/reference does not exist.  __graft_entry__.build() runs this whenever /root/reference is present.
Usage:  python oracle/build_ref.py
"""
import hashlib
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FILES = {"src/adapters/mona.py": "mona.py", "src/adapters/lora.py": "lora.py", "src/losses/losses.py": "losses.py",
         # the reference's training script, for tests/test_gpu_reference_loop.py: its own train() runs over this repository's
         # src.adapters / src.losses shims ("drops into src.models.biomedclip.finetune")
         "src/models/biomedclip/finetune.py": "finetune.py"}


def main():
    out = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(out, exist_ok=True)
    lines = []
    for rel, name in FILES.items():
        src = os.path.join(REF, rel)
        shutil.copyfile(src, os.path.join(out, name))
        lines.append(f"{hashlib.sha256(open(src, 'rb').read()).hexdigest()}  {rel}")
    open(os.path.join(out, "MANIFEST"), "w").write("\n".join(lines) + "\n")
    print("oracle/_ref:", ", ".join(FILES.values()))


def load(name):
    """Import oracle/_ref/<name>.py by path (None when the directory was not built)."""
    import importlib.util
    path = os.path.join(ROOT, "oracle", "_ref", name + ".py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("ngu_ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    main()
