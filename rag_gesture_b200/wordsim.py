"""Word similarity of the gesture-type retrieval rules (rag/utils.py:231-272).

The reference scores two gesture words with `word2vec_model.similarity` / `fasttext_model.similarity`
(rag/utils.py:231-236) and, on ANY exception, with `fuzz.partial_ratio(word1, word2) / 100`
(rag/utils.py:269-270).  Neither embedding model is defined anywhere in the reference, so the
shipped behaviour is the fall-back on every call; `word_similarity(..., model=None)` is exactly that.  A caller that
has an embedding model passes `model=callable(w1, w2) -> float` and gets the reference's multi-word averaging
around it (rag/utils.py:248-268) with the same fall-back.

`partial_ratio` restates fuzzywuzzy 0.18.0 (requirements.txt:14; the package is not in this image) on difflib --
the backend fuzzywuzzy itself uses when python-Levenshtein is absent, and the reference does not list it.
Known answers from the package's documentation are held in tests/test_retrieval_host.py.
"""
from difflib import SequenceMatcher


def partial_ratio(s1, s2):
    """fuzz.partial_ratio: best `ratio` of the shorter string against the equally long windows of the longer one
    that line up with a matching block; 0..100, rounded half-to-even as int(round(.)) does."""
    if s1 is None or s2 is None:
        return 0
    if s1 == s2:
        return 100
    if len(s1) == 0 or len(s2) == 0:
        return 0
    short, long_ = (s1, s2) if len(s1) <= len(s2) else (s2, s1)
    best = 0.0
    for a, b, _ in SequenceMatcher(None, short, long_).get_matching_blocks():
        lo = max(b - a, 0)
        r = SequenceMatcher(None, short, long_[lo:lo + len(short)]).ratio()
        if r > 0.995:
            return 100
        best = max(best, r)
    return int(round(100 * best))


def word_similarity(word1, word2, model=None):
    """get_word_similarity_score (rag/utils.py:239-272): `model` applied to the word pair, averaged over the words
    of multi-word entries; partial_ratio / 100 whenever the model is absent or raises."""
    try:
        if model is None:
            raise NameError("word2vec_model")              # what the shipped reference runs into
        a, b = word1.split(), word2.split()
        if len(a) > 1 and len(b) == 1:
            return sum(model(w, word2) for w in a) / len(a)
        if len(b) > 1 and len(a) == 1:
            return sum(model(word1, w) for w in b) / len(b)
        if len(a) > 1 and len(b) > 1:
            return sum(model(w1, w2) for w1 in a for w2 in b) / (len(a) * len(b))
        return model(word1, word2)
    except Exception:
        return partial_ratio(word1, word2) / 100
