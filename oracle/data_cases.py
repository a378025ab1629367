"""TEST INFRASTRUCTURE: the synthetic on-disk dataset used to pin the input pipeline (`phoneme-vqa_b200/data.py`)
against the reference's dataset classes.  Pure helpers — nothing here touches /root/reference; the generator script
`oracle/make_golden_data.py` does."""
import json
import os

import numpy as np

WORDS = ["cửa", "hàng", "bánh", "mì", "số", "12", "phở", "Hà", "Nội", "SALE", "50%", "đường", "Nguyễn", "Trãi", "café",
         "trà", "sữa", "MILK", "tea", "quán", "ăn", "ngon", "giá", "rẻ", "mở", "7h-22h", "wifi", "free", "ATM", "xăng"]
QUESTIONS = ["cửa hàng này bán gì ?", "số nhà là bao nhiêu", "biển hiệu màu gì", "quán mở cửa lúc mấy giờ ?",
             "tên đường là gì", "  có wifi không  ", "giá bao nhiêu", "đây là ở đâu"]
ANSWERS = ["bánh mì", "12", "màu đỏ", "7 giờ", "nguyễn trãi", "có", "năm mươi nghìn", "hà nội"]


def build_case(seed=7):
    rng = np.random.RandomState(seed)
    images = {}
    # image 3 has more OCR words than max_ocr_element and more sub-tokens than max_ocr_length; image 5 has none
    for image_id, n_words in [(1, 4), (2, 9), (3, 30), (5, 0), (8, 1), (13, 17)]:
        words = [WORDS[int(k)] for k in rng.randint(0, len(WORDS), size=n_words)]
        boxes = []
        for _ in range(n_words):
            x0, y0 = rng.uniform(0, 0.9, size=2)
            w, h = rng.uniform(0.001, 0.1, size=2)
            boxes.append([float(x0), float(y0), float(x0 + w), float(y0 + h)])
        feat = rng.standard_normal((1, 3, 4, 4)).astype(np.float32)
        images[image_id] = {"texts": words, "boxes": boxes, "feature": feat.tolist()}
    qa = []
    for k in range(14):
        image_id = [1, 2, 3, 5, 8, 13, 21][k % 7]          # 21 has no OCR file: dropped by the inner merge
        qa.append({"image_id": image_id, "question": QUESTIONS[k % len(QUESTIONS)], "answer": ANSWERS[(k * 3) % len(ANSWERS)],
                   "filename": f"{image_id}.jpg"})
    return {"images": images, "qa": qa,
            "params": {"max_ocr_element": 12, "max_ocr_length": 16, "max_input_length": 10, "max_output_length": 20}}


def write_case(case, root):
    """materialise the case in the reference's on-disk formats; returns (ocr_root, feature_root, qa_df)"""
    import pandas as pd
    ocr_root, feat_root = os.path.join(root, "ocr"), os.path.join(root, "features")
    os.makedirs(ocr_root, exist_ok=True)
    os.makedirs(feat_root, exist_ok=True)
    for image_id, im in case["images"].items():
        np.save(os.path.join(ocr_root, f"{image_id}.npy"),
                {"texts": im["texts"], "boxes": np.asarray(im["boxes"], dtype=np.float64).reshape(-1, 4)}, allow_pickle=True)
        # the int64 (QA table) x float64 (OCR table) inner merge keeps the left, integer ids: features are '3.npy'
        np.save(os.path.join(feat_root, f"{int(image_id)}.npy"), {"image": np.asarray(im["feature"], dtype=np.float32)},
                allow_pickle=True)
    qa_df = pd.DataFrame(case["qa"])[["image_id", "question", "answer", "filename"]]
    return ocr_root, feat_root, qa_df


def phoneme_tokenizer(case, root):
    """the product's 3-vocabulary tokenizer built from an annotation file of the case's questions and answers"""
    from importlib import import_module
    text = import_module("phoneme_vqa_b200.text")
    ann = os.path.join(root, "annotations.json")
    with open(ann, "w", encoding="utf-8") as f:
        json.dump({"annotations": [{"question": q["question"], "answers": [q["answer"]]} for q in case["qa"]]}, f,
                  ensure_ascii=False)
    return text.PhonemeTokenizer(vocab_path=None, annotation_paths=[ann], max_length=case["params"]["max_output_length"])


# ---------------------------------------------------------------------------------------------------------------
# SaL family case: OCR files double as OCR feature files (det/rec features), object files carry region features
# ---------------------------------------------------------------------------------------------------------------
OBJECTS = ["biển hiệu", "xe máy", "người", "cửa", "cây", "bàn ghế", "đèn", "tòa nhà"]
SAL_ANSWERS = ["bánh mì & trà", "12/5", "màu_đỏ", "7 giờ; tối", "a=b", "có", "năm mươi nghìn", "hà nội"]


def build_sal_case(seed=11):
    rng = np.random.RandomState(seed)
    images = {}
    for image_id, n_words, n_obj in [(1, 4, 3), (2, 9, 0), (3, 30, 12), (5, 0, 2), (8, 1, 1)]:
        words = [WORDS[int(k)] for k in rng.randint(0, len(WORDS), size=n_words)]
        xy = rng.uniform(0, 0.9, size=(n_words, 2)); wh = rng.uniform(0.001, 0.1, size=(n_words, 2))
        W, H = int(rng.randint(300, 900)), int(rng.randint(300, 900))
        oxy = rng.uniform(0, 0.8, size=(n_obj, 2)) * [W, H]; owh = rng.uniform(5, 60, size=(n_obj, 2))
        images[str(image_id)] = {
            "texts": words, "boxes": np.concatenate([xy, xy + wh], 1).tolist(),
            "det_features": rng.standard_normal((n_words, 4)).astype(np.float32).tolist(),
            "rec_features": rng.standard_normal((n_words, 2)).astype(np.float32).tolist(),
            "object_list": [OBJECTS[int(k)] for k in rng.randint(0, len(OBJECTS), size=n_obj)],
            "region_boxes": np.concatenate([oxy, oxy + owh], 1).tolist(), "height": H, "width": W,
            "region_features": rng.standard_normal((n_obj, 5)).astype(np.float32).tolist()}
    qa = []
    for k in range(11):
        image_id = [1, 2, 3, 5, 8, 21][k % 6]              # 21 has no files: dropped by the inner merges
        qa.append({"image_id": image_id, "question": QUESTIONS[k % len(QUESTIONS)], "answer": SAL_ANSWERS[(k * 3) % len(SAL_ANSWERS)],
                   "filename": f"{image_id}.jpg"})
    return {"images": images, "qa": qa,
            "params": {"ocr_hidden": 6, "obj_hidden": 5, "max_ocr_element": 12, "max_ocr_length": 20, "max_obj_element": 6,
                       "max_obj_length": 8, "max_input_length": 10, "max_output_length": 24}}


def write_sal_case(case, root):
    """-> (ocr_root, obj_root, qa_df) in the on-disk formats core/data/utils.py and PhonemeSaLDataset read"""
    import pandas as pd
    import torch
    ocr_root, obj_root = os.path.join(root, "sal_ocr"), os.path.join(root, "sal_obj")
    os.makedirs(ocr_root, exist_ok=True)
    os.makedirs(obj_root, exist_ok=True)
    for image_id, im in case["images"].items():
        n = len(im["texts"])
        np.save(os.path.join(ocr_root, f"{int(image_id)}.npy"),
                {"texts": im["texts"], "boxes": np.asarray(im["boxes"], dtype=np.float64).reshape(-1, 4),
                 "det_features": np.asarray(im["det_features"], dtype=np.float32).reshape(n, 4),
                 "rec_features": np.asarray(im["rec_features"], dtype=np.float32).reshape(n, 2)}, allow_pickle=True)
        m = len(im["object_list"])
        np.save(os.path.join(obj_root, f"{int(image_id)}.npy"),
                {"object_list": im["object_list"], "region_boxes": np.asarray(im["region_boxes"], dtype=np.float64).reshape(-1, 4),
                 "height": im["height"], "width": im["width"],
                 "region_features": torch.tensor(im["region_features"], dtype=torch.float32).reshape(m, 5)}, allow_pickle=True)
    qa_df = pd.DataFrame(case["qa"])[["image_id", "question", "answer", "filename"]]
    return ocr_root, obj_root, qa_df
