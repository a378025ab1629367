"""Deterministic word / sentence cases for the text-pipeline goldens (shared by the generator and the tests)."""
import random
import unicodedata

CURATED = """quyển giếng gì gìn giết giếc xin chào việt nam hỏa thủy thuở thỏa huề huế huệ nghiêng nghe ngủ ghế gà
kẻ cá quả quốc khuya khuỷu rượu hươu ngoằn ngoèo uống yêu yến ý ỷ y oanh oách oăm xoong boong quýt quỳnh
trường trưa mưa múa lúa của cửa thuế thuê tuyết chuyện nguyễn người ngoại toán hoàng hoa hoe hoẻn hoét
đường đẹp đi đâu được điện thoại biển báo cấm dừng lại bên trái phải trên dưới trong ngoài màu xanh đỏ vàng
tên cửa hàng là số mấy bao nhiêu giá tiền phở bún chả nem rán cơm tấm bánh mì cà phê sữa đá trà chanh
ăn uống ở ơ ư ô ê a ă â e i o u an ăn ân en ên in on ôn ơn un ưn
kia kìa mía chia phía nghĩa nghỉa quạ quế quy quỷ quít qui gia già giá giả giã giạ giô giơ giu giư
coca pepsi hello world abc xyz www http 123 2024 12h30 covid-19 wi-fi a4 b52 k+ ok no.1 (abc) "quote" it's
ATM KFC Việt NAM kfc atm đ d z f j w p pp ff zz oo ooo boo zoo queen king ring sing song long
ngh ng nh gh gi kh ph th tr ch q qu qua quo quu qa""".split()

RAW_SENTENCES = ["Bảng & Biển_báo #1 | ~test; a/b\\c = d", "  Xin   chào  ;Việt=Nam  ", "A&B", "no specials here",
                 "x_y#z|w~v"]


def _syllables():
    from itertools import product
    onsets = ["", "ngh", "tr", "th", "ph", "nh", "ng", "kh", "gi", "gh", "ch", "q", "đ", "x", "v", "t", "s", "r", "n",
              "m", "l", "k", "h", "g", "d", "c", "b"]
    rhymes = ("a ac ach ai am an ang anh ao ap at ay au ă ăc ăm ăn ăng ăp ăt â âc âm ân âng âp ât âu ây e ec em en eng eo "
              "ep et ê êch êm ên ênh êp êt êu i ia ich iêc iêm iên iêng iêp iêt iêu im in inh ip it iu o oa oac oach oai "
              "oam oan oang oanh oao oap oat oay oăc oăm oăn oăng oăt oc oe oen oeo oet oi om on ong ooc oong op ot ô ôc "
              "ôi ôm ôn ông ôp ôt ơ ơi ơm ơn ơp ơt u ua uân uâng uât uây uc uê uêch uênh ui um un ung uơ uôc uôi uôm uôn "
              "uông uôt up ut uy uya uych uyên uyêt uyn uynh uyp uyt uyu uach uai uan uang uanh uao uat uau uay uăc uăm "
              "uăn uăng uăp uăt uâc uoang ue uen ueo uet uên uêt uêu uơi ư ưa ưc ưi ưng ươc ươi ươm ươn ương ươp ươt ươu "
              "ưt ưu y yêm yên yêng yêt yêu").split()
    marks = ["", "̀", "́", "̃", "̉", "̣"]
    vowels = set("aăâeêioôơuưy")
    out = []
    for o, r, m in product(onsets, rhymes, marks):
        if m:
            k = next(i for i, ch in enumerate(r) if ch in vowels)
            r2 = r[: k + 1] + m + r[k + 1:]
        else:
            r2 = r
        out.append(unicodedata.normalize("NFC", o + r2))
    return out


def _random_words(n=6000, seed=20240517):
    rng = random.Random(seed)
    alphabet = "abcdeghiklmnopqrstuvxyăâêôơưđáàảãạếềểễệóòỏõọúùủũụíìỉĩịýỳỷỹỵ0123456789-.,'"
    return ["".join(rng.choice(alphabet) for _ in range(rng.randint(1, 7))) for _ in range(n)]


def all_words():
    rw = _random_words()
    return CURATED + rw[:400] + _syllables() + rw[400:]


def sentences():
    base = ["xin chào việt nam", "biển báo cấm dừng xe", "tên cửa hàng là gì", "số điện thoại 0912345678",
            "quán cà phê sữa đá", "giá tiền là 25.000 đ", "màu xanh lá cây", "coca cola", "phở bò tái chín",
            "nguyễn văn a", "đường trường chinh", "khuya", "wi-fi free", "a", "", "quyển sách tiếng việt",
            "hello world 2024", "giếng nước", "gì", "yêu"]
    rng = random.Random(7)
    words = [w for w in CURATED if w.islower() and w.isalpha()]
    extra = [" ".join(rng.choice(words) for _ in range(rng.randint(1, 12))) for _ in range(20)]
    return base + extra


def annotations():
    return {"annotations": [{"question": "Tên cửa hàng là gì ?", "answers": ["phở bò"]},
                            {"question": "số điện thoại wi-fi ?", "answers": ["0912 abc"]},
                            {"question": "quyển sách màu gì", "answers": "màu xanh"},
                            {"question": "giếng ở đâu", "answers": ["bên trái coca"]}]}
