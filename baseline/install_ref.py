#!/usr/bin/env python
"""Install the UNMODIFIED reference under baseline/_ref/ so that it travels to the GPU box.

    python baseline/install_ref.py            # build container only: needs /root/reference

cujoramirez/QA-ViT has no setup.py / pyproject: it is a set of stand-alone model scripts, so the
"install" of the bench contract (`pip install --target baseline/_ref /root/reference`) degenerates to copying the
model scripts byte for byte.  baseline/_ref/ is git-ignored (it never enters the history) but NOT gpurun-ignored, so
`bench.py --impl reference`, the gpu_eager_baseline leg and tests/test_gpu_live_reference.py can import the live
reference on the B200 box, where /root/reference does not exist.  A sha256 manifest is written next to the copies;
nothing is edited (flash-attn selection etc. happen at run time through the modules' own globals).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("QAVIT_REFERENCE", "/root/reference")
# the model scripts only (the reference's train/eval drivers import matplotlib / seaborn / torchvision datasets)
FILES = ["HQAViT_CIFAR100.py", "HQAViTv2_CIFAR100.py", "HQAViT_IN_Tiny.py", "QAViT.py", "QAViTv2.py", "QAViTv2_CIFAR100.py",
         "QAViTV2_EXTREME.py", "HQAViT_C100_Finetune.py", "HQAViT_Tiny_stl10.py", "HQAViT_Tiny_Cifar10.py", "test_hqa.py"]


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"install_ref: {SRC} not present (GPU box?) -- keeping whatever baseline/_ref already holds")
        return 0 if os.path.isdir(DST) else 1
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        s = os.path.join(SRC, f)
        if not os.path.exists(s):
            continue
        d = os.path.join(DST, f)
        shutil.copyfile(s, d)
        os.chmod(d, 0o644)
        manifest[f] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    print(f"install_ref: {len(manifest)} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
