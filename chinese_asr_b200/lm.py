"""Second-pass language model with device-resident tables (replaces kenlm at model.py:749-763).

The reference calls kenlm.LanguageModel(lm_path).score(' '.join(chars), bos=True).  KenLM and its
binary model are not available here, so the B200 path takes a back-off n-gram (order <= 3) in ARPA
semantics: log10 probabilities, back-off weights, <s> context when bos, </s> appended when eos,
a word the LM does not know scores as <unk> (kenlm maps it to its <unk> index before every lookup:
dict.pkl tokens missing from an ARPA file go through `id_map`; n-grams over words that dict.pkl does
not have can never be asked for and are dropped).  Scoring of the finished hypotheses runs on the GPU (csrc/decoder.cu); this class
only builds / uploads the tables.  `.score()` is provided for API compatibility and evaluates the
same tables through the device kernel (asr_lm_score) - there is no host scoring path."""
import numpy as np

SPACE_ID = 781      # dict.pkl's literal ' ' token: vanishes under str.split()
UNK = 3


def _hash_slot(key, cap):
    m = (1 << 64) - 1
    z = (key + 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    z = z ^ (z >> 31)
    return int(z & (cap - 1))


def _build_table(items, nvals):
    cap = 1
    while cap < 2 * len(items) + 8:
        cap *= 2
    keys = np.full(cap, -1, dtype=np.int64)
    vals = np.zeros((cap, nvals), dtype=np.float32)
    for k, v in items:
        s = _hash_slot(k, cap)
        while keys[s] != -1:
            s = (s + 1) & (cap - 1)
        keys[s] = k
        vals[s] = v
    return keys, vals


class NGramLM:
    def __init__(self, tables, word2int=None, skip_id=SPACE_ID):
        """tables: dict with uni_logp[V], uni_bo[V], bi_keys, bi_vals[cap,2], tri_keys, tri_vals."""
        self.t = {k: np.ascontiguousarray(v) for k, v in tables.items()}
        self.t["bi_vals"] = self.t["bi_vals"].reshape(-1, 2).astype(np.float32)
        self.t["tri_vals"] = self.t["tri_vals"].reshape(-1).astype(np.float32)
        self.vocab = int(self.t["uni_logp"].shape[0])
        self.word2int = word2int
        self.skip_id = skip_id
        self._engine = None
        if "id_map" in self.t:
            self.t["id_map"] = self.t["id_map"].astype(np.int32)

    def tables(self):
        return self.t

    @classmethod
    def from_arpa(cls, path, word2int, skip_id=SPACE_ID):
        """Load an ARPA file (order <= 3) whose words are dict.pkl strings."""
        V = len(word2int)
        uni_logp = np.full(V, -99.0, dtype=np.float32)
        uni_bo = np.zeros(V, dtype=np.float32)
        bi, tri = [], []
        seen = np.zeros(V, dtype=bool)
        order = 0
        with open(path, encoding='utf-8') as f:
            for line in f:
                line = line.rstrip('\n')
                if line.startswith('\\') and line.endswith('-grams:'):
                    order = int(line[1])
                    continue
                if not line or line.startswith('\\') or line.startswith('ngram '):
                    continue
                if order == 0:
                    continue
                if '\t' in line:
                    parts = line.split('\t')
                    words = parts[1].split(' ')
                    bo = float(parts[2]) if len(parts) > 2 else 0.0
                else:
                    parts = line.split(' ')
                    words = parts[1:1 + order]
                    bo = float(parts[1 + order]) if len(parts) > 1 + order else 0.0
                if len(words) != order:
                    continue
                lp = float(parts[0])
                # "<unk>" itself is dict.pkl id 3; any other word dict.pkl lacks can never be produced by the
                # decoder, so n-grams containing it are unreachable (mapping them onto <unk> would overwrite
                # <unk>'s own entries)
                if any(w not in word2int for w in words):
                    continue
                ids = [word2int[w] for w in words]
                if order == 1:
                    seen[ids[0]] = True
                    uni_logp[ids[0]] = lp
                    uni_bo[ids[0]] = bo
                elif order == 2:
                    bi.append((ids[0] * V + ids[1], (lp, bo)))
                elif order == 3:
                    tri.append(((ids[0] * V + ids[1]) * V + ids[2], (lp,)))
        # later duplicates of an n-gram must not shadow the first (kenlm rejects them; first one wins here)
        bi = list({k: v for k, v in reversed(bi)}.items())
        tri = list({k: v for k, v in reversed(tri)}.items())
        bk, bv = _build_table(bi, 2)
        tk, tv = _build_table(tri, 1)
        # dict.pkl tokens the ARPA file has no unigram for are out-of-vocabulary for the LM: kenlm looks them
        # up as <unk>, in every position of an n-gram
        id_map = np.where(seen, np.arange(V), UNK).astype(np.int32)
        return cls({"uni_logp": uni_logp, "uni_bo": uni_bo, "bi_keys": bk, "bi_vals": bv,
                    "tri_keys": tk, "tri_vals": tv.reshape(-1), "id_map": id_map}, word2int, skip_id)

    def bind(self, engine):
        self._engine = engine

    def score(self, sentence, bos=True, eos=True):
        """kenlm-style total log10 probability, evaluated by the device kernel."""
        if not (bos and eos):
            raise NotImplementedError("the device scorer implements bos=True, eos=True (model.py:755)")
        if self._engine is None or self.word2int is None:
            raise RuntimeError("NGramLM.score needs bind(model) and word2int")
        ids = [self.word2int.get(w, UNK) for w in sentence.split()]
        return float(self._engine.lm_score([ids])[0])
