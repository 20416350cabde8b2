"""Real-data fixture for the image contract P2 (TEST INFRASTRUCTURE, see oracle/__init__.py).

Runs the reference's own image preprocessing -- ``Image.open(...).convert('RGB')`` -> ``Resize((128, 128))`` ->
``ToTensor()`` (/root/reference/Descriptors/multi_input_data_preprocess_maccs_opt.py:52-67) -- over the 1 058 depictions
the reference ships (Descriptors/img_output/{NO}.png, drawn by convert_smiles_2_img.py) and stores the result as uint8
CHW: ``ToTensor`` of an 8-bit image is exactly ``u / 255`` in float32 (asserted below), so the uint8 array IS the
reference's ``Image_Features`` without loss.  Labels are the ``logBB`` column of B3DB/B3DB/B3DB_regression.tsv, row
order = ``NO.`` order.  (MACCS / Morgan bits need RDKit, which this image does not have: fingerprints stay synthetic.)

    python oracle/make_real_fixture.py        # needs /root/reference; writes tests/golden/b3db_depictions_u8.npz
"""
from __future__ import annotations

import os

import numpy as np

REF = os.environ.get("BBBP_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "b3db_depictions_u8.npz")


def main() -> None:
    import pandas as pd
    from PIL import Image
    import torchvision.transforms as transforms

    table = pd.read_csv(os.path.join(REF, "B3DB", "B3DB", "B3DB_regression.tsv"), sep="\t")
    transform = transforms.Compose([transforms.Resize((128, 128)), transforms.ToTensor()])
    imgs = []
    for no in table["NO."]:
        img = Image.open(os.path.join(REF, "Descriptors", "img_output", f"{no}.png")).convert("RGB")
        x = transform(img).numpy()                                   # the reference's Image_Features (before flatten)
        u = np.round(x * 255.0).astype(np.uint8)
        assert np.array_equal(u.astype(np.float32) / np.float32(255.0), x)
        imgs.append(u)
    np.savez_compressed(OUT, img=np.stack(imgs), logBB=table["logBB"].to_numpy(np.float32),
                        no=table["NO."].to_numpy(np.int32))
    print(OUT, os.path.getsize(OUT), "bytes;", len(imgs), "molecules")


if __name__ == "__main__":
    main()
