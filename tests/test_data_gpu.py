"""Device side of the input pipeline: pinned rotating buffers + host->device copies on a side stream."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pinned_loader_delivers_identical_batches_on_device(tmp_path):
    from oracle.data_cases import phoneme_tokenizer, write_case
    from oracle.stub_tokenizer import StubT5Tokenizer
    from phoneme_vqa_b200 import data
    with open(os.path.join(ROOT, "tests", "golden", "data_phonemelatr.json"), encoding="utf-8") as f:
        case = json.load(f)["case"]
    ocr_root, feat_root, qa_df = write_case(case, str(tmp_path))
    ocr_df = data.textlayout_ocr_adapt(ocr_root).sort_values("image_id").reset_index(drop=True)
    p = case["params"]
    ds = data.PhonemeLaTrDataset(qa_df, ocr_df, StubT5Tokenizer(), phoneme_tokenizer(case, str(tmp_path)), feat_root,
                                 max_ocr_element=p["max_ocr_element"], max_ocr_length=p["max_ocr_length"],
                                 max_input_length=p["max_input_length"], max_output_length=p["max_output_length"])
    ds.pack_features(str(tmp_path / "pack.npy"))
    g1, g2 = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    host = list(data.PinnedBatchLoader(ds, 5, shuffle=True, generator=g1))
    loader = data.PinnedBatchLoader(ds, 5, shuffle=True, generator=g2, device="cuda:0", prefetch=2)
    n = 0
    acc = torch.zeros((), device="cuda:0")
    for epoch in range(3):                       # buffers are recycled several times
        if epoch:
            loader.generator = torch.Generator().manual_seed(5)
        for x, y in zip(loader, host):
            assert all(v.is_cuda for v in x.values())
            for k in y:
                assert torch.equal(x[k].cpu(), y[k]), (epoch, k)
            acc += x["pixel_values"].sum()       # consumer work on the current stream
            n += 1
    assert n == 3 * len(host)
    assert loader._buffers[0]["pixel_values"].is_pinned()
    torch.cuda.synchronize()
