"""Stage the reference's hot-path modules for the GPU box: copies the 17 files that oracle/refshim.load()
imports (the denoiser, attention, StylizationBlock, GaussianDiffusion / SpacedDiffusion, the retrieval
functions, the architecture) UNMODIFIED from /root/reference into oracle/_ref/ (git-ignored, NOT
gpurun-ignored), so `bench.py --impl reference` and the `cpu_baseline` leg can time the reference's own
code on the box's host cores.  Run in the build container (where /root/reference exists):

    python oracle/stage_ref.py

__graft_entry__.build() calls stage() when the reference tree is present.  Nothing under oracle/_ref/ is
ever committed, and nothing in the product path reads it."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("RG_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [
    "mogen/models/architectures/base_architecture.py",
    "mogen/models/architectures/diffusion_architecture.py",
    "mogen/models/attentions/efficient_attention.py",
    "mogen/models/builder.py",
    "mogen/models/losses/mse_loss.py",
    "mogen/models/losses/utils.py",
    "mogen/models/transformers/diffusion_transformer.py",
    "mogen/models/transformers/gesture_vae.py",
    "mogen/models/transformers/rag/discourse_retrieval.py",
    "mogen/models/transformers/rag/gesture_type_retrieval.py",
    "mogen/models/transformers/rag/llm_retrieval.py",
    "mogen/models/transformers/rag/utils.py",
    "mogen/models/transformers/raggesture.py",
    "mogen/models/utils/detr_utils.py",
    "mogen/models/utils/gaussian_diffusion.py",
    "mogen/models/utils/rotation_conversions.py",
    "mogen/models/utils/stylization_block.py",
]


def stage(verbose=False):
    """Returns the number of files staged (0 when the reference tree is absent: the GPU box)."""
    if not os.path.isdir(os.path.join(SRC, "mogen")):
        return 0
    n = 0
    for f in FILES:
        src, dst = os.path.join(SRC, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src) or \
                os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)
        n += 1
    if verbose:
        print(f"staged {n} reference files under {DST}")
    return n


if __name__ == "__main__":
    if not stage(verbose=True):
        sys.exit(f"reference tree not found at {SRC}")
