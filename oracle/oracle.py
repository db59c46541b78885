"""ctypes binding of the ORACLE (CPU restatement of lofcz/Infidex's search path).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product (infidex_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_DIR, "libifx_oracle.so")


def build(force=False):
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith((".cpp", ".hpp", ".inc"))]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _DIR, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.ifxo_create.restype = C.c_void_p
        _lib.ifxo_avgdl.restype = C.c_float
        _lib.ifxo_word_idf.restype = C.c_float
    return _lib


def u16(s):
    return np.frombuffer(s.encode("utf-16-le", "surrogatepass"), dtype=np.uint16).copy() if s else np.zeros(0, np.uint16)


def from_u16(a):
    return np.asarray(a, dtype=np.uint16).tobytes().decode("utf-16-le", "surrogatepass")


def pack_strings(strs):
    enc = [s.encode("utf-16-le", "surrogatepass") for s in strs]
    offs = np.zeros(len(enc) + 1, np.int64)
    np.cumsum([len(e) // 2 for e in enc], out=offs[1:])
    blob = np.frombuffer(b"".join(enc), dtype=np.uint16).copy() if offs[-1] else np.zeros(1, np.uint16)
    return blob, offs


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Field:
    def __init__(self, name, weight=1, indexable=True, filterable=False, facetable=False):
        self.name, self.weight, self.indexable, self.filterable, self.facetable = name, weight, indexable, filterable, facetable


def fst_query(terms, query, cap=10, prefix=False):
    """FstIndex built from {term: output}: MatchWithinEditDistance1 (or GetByPrefix) -> (return value, outputs written)."""
    keys = list(terms); blob = u16("".join(keys)); off = np.zeros(len(keys) + 1, np.int32); off[1:] = np.cumsum([len(k) for k in keys])
    outs = np.array([terms[k] for k in keys], np.int32); q = u16(query); out = np.full(max(cap, 1), -1, np.int32)
    n = lib().ifxo_fst_query(_p(blob), _p(off), _p(outs), len(keys), int(prefix), _p(q), len(q), _p(out), cap)
    return n, [int(x) for x in out[:min(n, cap)]]


class OracleEngine:
    """SearchEngine.CreateDefault() restated (config 400)."""

    def __init__(self, schema=None):
        self.h = C.c_void_p(lib().ifxo_create())
        self.schema = schema or [Field("content", 1)]
        blob, offs = pack_strings([f.name for f in self.schema])
        offs32 = offs.astype(np.int32)
        w = np.array([f.weight for f in self.schema], np.int32)
        fl = np.array([(1 if f.indexable else 0) | (2 if f.filterable else 0) | (4 if f.facetable else 0) for f in self.schema], np.int32)
        lib().ifxo_set_schema(self.h, len(self.schema), _p(blob), _p(offs32), _p(w), _p(fl))

    def __del__(self):
        try:
            lib().ifxo_destroy(self.h)
        except Exception:
            pass

    def index_columns(self, keys, columns):
        """columns: list (per schema field) of list[str] | np.int64 array | np.float64 array."""
        n = len(keys)
        keys = np.ascontiguousarray(keys, np.int64)
        kinds = np.zeros(len(columns), np.int32)
        keep, cols, offs = [], (C.c_void_p * len(columns))(), (C.c_void_p * len(columns))()
        for i, col in enumerate(columns):
            if isinstance(col, np.ndarray) and col.dtype.kind in "iu":
                a = np.ascontiguousarray(col, np.int64); kinds[i] = 2; cols[i] = a.ctypes.data; keep.append(a)
            elif isinstance(col, np.ndarray) and col.dtype.kind == "f":
                a = np.ascontiguousarray(col, np.float64); kinds[i] = 3; cols[i] = a.ctypes.data; keep.append(a)
            elif isinstance(col, tuple):
                blob, o = col; kinds[i] = 1; cols[i] = blob.ctypes.data; offs[i] = o.ctypes.data; keep += [blob, o]
            else:
                blob, o = pack_strings(col); kinds[i] = 1; cols[i] = blob.ctypes.data; offs[i] = o.ctypes.data; keep += [blob, o]
        lib().ifxo_add_docs(self.h, n, _p(keys), _p(kinds), cols, offs)
        lib().ifxo_build(self.h)

    def load_image(self, image_ptr):
        """bench.py only: take the index state from a flattened ifx_index_image (pointer from SearchEngine.image_ptr()); see Index::load_image."""
        lib().ifxo_load_image(self.h, C.c_void_p(image_ptr))

    def index_texts(self, texts, keys=None):
        keys = np.arange(len(texts), dtype=np.int64) if keys is None else keys
        self.index_columns(keys, [list(texts)])

    def search(self, text, max_results=10, depth=500, coverage=True, filter_bytes=None, facets=False, cap=None):
        cap = cap or max(max_results, 1)
        q = u16(text)
        keys = np.zeros(cap, np.int64); scores = np.zeros(cap, np.float32); ties = np.zeros(cap, np.uint8)
        n = C.c_int(0); total = C.c_int(0); fb = C.create_string_buffer(1 << 16); fl = C.c_int(0)
        fbytes = np.frombuffer(filter_bytes, np.uint8).copy() if filter_bytes else None
        st = lib().ifxo_search(self.h, _p(q), len(q), max_results, depth, int(coverage),
                               _p(fbytes) if fbytes is not None else None, len(fbytes) if fbytes is not None else 0, int(facets),
                               _p(keys), _p(scores), _p(ties), cap, C.byref(n), C.byref(total), fb, len(fb), C.byref(fl))
        facet_list = []
        if facets:
            for line in fb.raw[: fl.value].decode("utf-8").splitlines():
                f, v, c = line.split("\t"); facet_list.append((f, v, int(c)))
        return {"status": st, "keys": keys[: n.value].tolist(), "scores": scores[: n.value].copy(), "ties": ties[: n.value].tolist(),
                "total": total.value, "facets": facet_list}

    def stage1(self, text, depth=500):
        q = u16(text)
        keys = np.zeros(depth, np.int64); scores = np.zeros(depth, np.float32); n = C.c_int(0); stats = np.zeros(5, np.int64)
        st = lib().ifxo_stage1(self.h, _p(q), len(q), depth, _p(keys), _p(scores), depth, C.byref(n), _p(stats))
        return {"status": st, "keys": keys[: n.value].copy(), "scores": scores[: n.value].copy(),
                "path": int(stats[0]), "candidates": int(stats[1]), "streamed": int(stats[2]), "n_terms": int(stats[3]), "n_fuzzy": int(stats[4])}

    def search_batch(self, queries, max_results=10, depth=500, coverage=True, filter_bytes=None, threads=1):
        blob, offs = pack_strings(queries); nq = len(queries); cap = max_results
        keys = np.zeros((nq, cap), np.int64); scores = np.zeros((nq, cap), np.float32); ties = np.zeros((nq, cap), np.uint8)
        ns = np.zeros(nq, np.int32); status = np.zeros(nq, np.int32)
        fbytes = np.frombuffer(filter_bytes, np.uint8).copy() if filter_bytes else None
        lib().ifxo_search_batch(self.h, _p(blob), _p(offs), nq, max_results, depth, int(coverage),
                                _p(fbytes) if fbytes is not None else None, len(fbytes) if fbytes is not None else 0, threads,
                                _p(keys), _p(scores), _p(ties), cap, _p(ns), _p(status))
        return keys, scores, ties, ns, status

    # ---- introspection
    def num_terms(self):
        return lib().ifxo_num_terms(self.h)

    def term(self, t):
        buf = np.zeros(256, np.uint16); n = lib().ifxo_term_text(self.h, t, _p(buf), 256)
        df = lib().ifxo_term_df(self.h, t); npost = lib().ifxo_term_npost(self.h, t)
        docs = np.zeros(max(npost, 1), np.int32); w = np.zeros(max(npost, 1), np.uint8)
        if npost: lib().ifxo_term_postings(self.h, t, _p(docs), _p(w))
        return from_u16(buf[:n]), df, docs[:npost], w[:npost]

    def lookup_term(self, s):
        a = u16(s); return lib().ifxo_lookup_term(self.h, _p(a), len(a))

    def doc_lens(self):
        out = np.zeros(lib().ifxo_num_docs(self.h), np.float32); lib().ifxo_doc_lens(self.h, _p(out)); return out

    def avgdl(self):
        return float(lib().ifxo_avgdl(self.h))

    def wm_lookup(self, word, affix=False):
        a = u16(word); out = np.zeros(4096, np.int32); n = lib().ifxo_wm_lookup(self.h, int(affix), _p(a), len(a), _p(out), len(out))
        return out[:max(n, 0)].tolist()

    def set_word_idf(self, cache):
        keys = list(cache); blob = u16("".join(keys)); off = np.zeros(len(keys) + 1, np.int32); off[1:] = np.cumsum([len(k) for k in keys])
        lib().ifxo_set_word_idf(self.h, _p(blob), _p(off), _p(np.array([cache[k] for k in keys], np.float32)), len(keys))

    def coverage(self, query, doc, lcs=0.0, bm25=0.0):
        q, d = u16(query), u16(doc); out = np.zeros(4, np.int32)
        lib().ifxo_coverage(self.h, _p(q), len(q), _p(d), len(d), C.c_double(lcs), C.c_float(bm25), _p(out))
        return {"coverage": int(out[0]), "word_hits": int(out[1]), "score": float(out[2:3].view(np.float32)[0]), "tie": int(out[3])}

    def filter_eval(self, filter_bytes, doc):
        fb = np.frombuffer(filter_bytes, np.uint8).copy()
        return lib().ifxo_filter_eval(self.h, _p(fb), len(fb), doc)


def levenshtein(a, b, max_errors=2**31 - 1, ignore_case=False):
    x, y = u16(a), u16(b); return lib().ifxo_levenshtein(_p(x), len(x), _p(y), len(y), max_errors, int(ignore_case))


def damerau(a, b, maxd, ignore_case=False):
    x, y = u16(a), u16(b); return lib().ifxo_damerau(_p(x), len(x), _p(y), len(y), maxd, int(ignore_case))
