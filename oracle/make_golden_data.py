"""TEST INFRASTRUCTURE: golden vectors for the input pipeline (`phoneme-vqa_b200/data.py`).

Runs in the BUILD container only (needs /root/reference): builds a small on-disk dataset in the reference's formats
(OCR '<id>.npy' pickles, feature '<id>.npy' pickles, a QA table), instantiates the REFERENCE classes
`core.data.textlayout_ocr_adapt` and `core.data.PhonemeLaTrDataset` on it with `oracle/stub_tokenizer.StubT5Tokenizer`
(the real T5 sentencepiece model is not available offline) and the product's 3-vocabulary `PhonemeTokenizer` (the
reference snapshot does not define one), and records every item.  The fixture also carries the raw case
(questions, answers, OCR words/boxes, feature arrays) so the test can rebuild the files anywhere.

usage: python oracle/make_golden_data.py   -> tests/golden/data_phonemelatr.json
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle.stub_tokenizer import StubFlatTokenizer, StubT5Tokenizer  # noqa: E402

from oracle.data_cases import build_case, build_sal_case, phoneme_tokenizer, write_case, write_sal_case  # noqa: E402


def dump_items(ds):
    items = []
    for i in range(len(ds)):
        it = ds[i]
        items.append({k: {"dtype": str(v.dtype), "shape": list(v.shape), "data": v.flatten().tolist()} for k, v in it.items()})
    return items


def main():
    from core.data import (CustomizedLaTrDataset, CustomizedPreSTUDataset, LaTrDataset, PhonemeLaTrDataset,      # the real
                           PhonemePreSTUDataset, textlayout_ocr_adapt)                                           # reference
    case = build_case()
    case["images"] = {str(k): v for k, v in case["images"].items()}
    with tempfile.TemporaryDirectory() as tmp:
        ocr_root, feat_root, qa_df = write_case(case, tmp)
        ocr_df = textlayout_ocr_adapt(ocr_root).sort_values("image_id").reset_index(drop=True)
        p = case["params"]
        kw = dict(max_ocr_element=p["max_ocr_element"], max_ocr_length=p["max_ocr_length"],
                  max_input_length=p["max_input_length"], max_output_length=p["max_output_length"])
        ds = PhonemeLaTrDataset(qa_df, ocr_df, StubT5Tokenizer(), phoneme_tokenizer(case, tmp), feat_root, **kw)
        items = dump_items(ds)
        prestu = dump_items(PhonemePreSTUDataset(qa_df, ocr_df, StubT5Tokenizer(), phoneme_tokenizer(case, tmp), feat_root, **kw))
        variants = {
            "LaTrDataset": dump_items(LaTrDataset(qa_df, ocr_df, StubT5Tokenizer(), feat_root, **kw)),
            "CustomizedLaTrDataset": dump_items(CustomizedLaTrDataset(qa_df, ocr_df, StubT5Tokenizer(), StubFlatTokenizer(), feat_root, **kw)),
            # PreSTUDataset cannot be instantiated in the snapshot: data_processing calls create_properties but the
            # class defines create_features (core/data/PreSTUDataset.py:69,87) -> no golden, parity unpinned
            "CustomizedPreSTUDataset": dump_items(CustomizedPreSTUDataset(qa_df, ocr_df, StubT5Tokenizer(), StubFlatTokenizer(), feat_root, **kw)),
        }
        ocr_rows = [{"image_id": float(r.image_id), "texts": list(r.texts), "bboxes": [list(map(float, b)) for b in r.bboxes]}
                    for r in ocr_df.itertuples()]
    out = {"case": case, "n_items": len(items), "image_ids": [float(x) for x in ds.data["image_id"]], "items": items,
           "items_prestu": prestu, "ocr_table": ocr_rows}
    vpath = os.path.join(ROOT, "tests", "golden", "data_variants.json")
    with open(vpath, "w", encoding="utf-8") as f:
        json.dump(variants, f, ensure_ascii=False)
    print("wrote", vpath, {k: len(v) for k, v in variants.items()}, "bytes", os.path.getsize(vpath))
    path = os.path.join(ROOT, "tests", "golden", "data_phonemelatr.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False)
    print("wrote", path, "items", len(items), "+", len(prestu), "bytes", os.path.getsize(path))
    main_sal()


def main_sal():
    """PhonemeSaLDataset of the real reference on the SaL case -> tests/golden/data_phonemesal.json"""
    from importlib import import_module
    from core.data import CustomizedSaLDataset, PhonemeSaLDataset, SaLDataset, textlayout_obj_adapt, textlayout_ocr_adapt
    text = import_module("phoneme_vqa_b200.text")
    case = build_sal_case()
    with tempfile.TemporaryDirectory() as tmp:
        ocr_root, obj_root, qa_df = write_sal_case(case, tmp)
        ocr_df = textlayout_ocr_adapt(ocr_root, h_scale=1, w_scale=1).sort_values("image_id").reset_index(drop=True)
        obj_df = textlayout_obj_adapt(obj_root, h_scale=1, w_scale=1).sort_values("image_id").reset_index(drop=True)
        p = case["params"]
        ds = PhonemeSaLDataset(qa_df, ocr_df, obj_df, StubT5Tokenizer(), text.FlatPhonemeTokenizer(), ocr_root, obj_root,
                               p["ocr_hidden"], p["obj_hidden"], max_ocr_element=p["max_ocr_element"],
                               max_ocr_length=p["max_ocr_length"], max_obj_element=p["max_obj_element"],
                               max_obj_length=p["max_obj_length"], max_input_length=p["max_input_length"],
                               max_output_length=p["max_output_length"])
        items = dump_items(ds)
        skw = dict(max_ocr_element=p["max_ocr_element"], max_ocr_length=p["max_ocr_length"], max_obj_element=p["max_obj_element"],
                   max_obj_length=p["max_obj_length"], max_input_length=p["max_input_length"],
                   max_output_length=p["max_output_length"])
        sal_variants = {
            "SaLDataset": dump_items(SaLDataset(qa_df, ocr_df, obj_df, StubT5Tokenizer(), ocr_root, obj_root,
                                                p["ocr_hidden"], p["obj_hidden"], **skw)),
            "CustomizedSaLDataset": dump_items(CustomizedSaLDataset(qa_df, ocr_df, obj_df, StubT5Tokenizer(), StubFlatTokenizer(),
                                                                    ocr_root, obj_root, p["ocr_hidden"], p["obj_hidden"], **skw)),
        }
        obj_rows = [{"image_id": float(r.image_id), "obj_labels": list(r.obj_labels),
                     "obj_bboxes": [list(map(float, b)) for b in r.obj_bboxes]} for r in obj_df.itertuples()]
    out = {"case": case, "n_items": len(items), "image_ids": [float(x) for x in ds.data["image_id"]], "items": items,
           "obj_table": obj_rows, "variants": sal_variants}
    path = os.path.join(ROOT, "tests", "golden", "data_phonemesal.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False)
    print("wrote", path, "items", len(items), "bytes", os.path.getsize(path))


if __name__ == "__main__":
    main()
