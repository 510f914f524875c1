"""GPU: SURVEY.md §8f rank 4 — decoded uint8 images -> bf16 [N, 3, 224, 224] on the device, bit-exact against the
reference's own transform (torchvision Resize(BICUBIC) on PIL -> ToTensor -> Normalize, mimicit_dataset.py:132-143),
and the whole-batch path (collate + H2D + label masking) in the layout `forward_pass` consumes."""
import numpy as np
import pytest
import torch

from oracle import preprocess as PP
from test_data_cpu import SIZES, _img

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference_transform():
    from torchvision import transforms
    return transforms.Compose([transforms.Resize((224, 224), interpolation=transforms.InterpolationMode.BICUBIC),
                               transforms.ToTensor(), transforms.Normalize(mean=PP.FLAMINGO_MEAN, std=PP.FLAMINGO_STD)])


def test_device_preprocess_bit_exact_vs_torchvision():
    from PIL import Image
    from otter_b200.data import ImagePreprocessor
    imgs = [_img(h, w, seed=h + w) for h, w in SIZES]
    t = _reference_transform()
    ref = torch.stack([t(Image.fromarray(im)) for im in imgs])                     # fp32 [N, 3, 224, 224]
    got32 = ImagePreprocessor(224, device=DEV, out_dtype=torch.float32)(imgs)
    assert torch.equal(got32.cpu(), ref)                                            # bit-exact fp32
    got = ImagePreprocessor(224, device=DEV)(imgs)
    assert got.dtype == torch.bfloat16 and torch.equal(got.cpu(), ref.to(torch.bfloat16))   # == images.to(autocast_type)
    # PIL inputs, a second call (cached coefficient tables), ragged batch of one
    pre = ImagePreprocessor(224, device=DEV)
    one = pre([Image.fromarray(imgs[3])])
    assert torch.equal(one.cpu(), ref[3:4].to(torch.bfloat16))
    assert torch.equal(pre(imgs[:2]).cpu(), ref[:2].to(torch.bfloat16))


def test_device_batcher_layout_and_labels():
    from otter_b200.data import DeviceBatcher
    from oracle.restatement import label_mask_np
    EOS, ANS, EOC, PAD = 2, 7, 8, 1
    g = torch.Generator().manual_seed(0)
    samples = []
    for n in (12, 20, 7):
        ids = torch.randint(9, 40, (n,), generator=g)
        ids[2], ids[-2], ids[-1] = ANS, EOC, EOS
        samples.append({"source": ids, "text_mask": torch.ones(n, dtype=torch.int64),
                        "images": [_img(60 + n, 90, seed=n), _img(224, 224, seed=n + 1)]})
    batch = DeviceBatcher(PAD, EOS, answer_token_id=ANS, endofchunk_token_id=EOC, device=DEV)(samples)
    ni = batch["net_input"]
    assert ni["patch_images"].shape == (3, 1, 2, 3, 224, 224) and ni["patch_images"].dtype == torch.bfloat16
    assert ni["input_ids"].shape == (3, 20) and ni["input_ids"].is_cuda
    assert ni["input_ids"][0, 12:].eq(PAD).all() and ni["attention_masks"][2, 7:].eq(0).all()
    want = label_mask_np(ni["input_ids"].cpu().numpy(), EOS, ANS, EOC)
    assert np.array_equal(batch["labels"].cpu().numpy(), want)
    ref = torch.from_numpy(PP.patch_resize_transform(samples[1]["images"][0])).to(torch.bfloat16)
    assert torch.equal(ni["patch_images"][1, 0, 0].cpu(), ref)
