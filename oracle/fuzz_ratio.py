"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): stand-in for `fuzzywuzzy.fuzz.partial_ratio`.

The reference's gesture-type rules call `fuzz.partial_ratio` (rag/utils.py:270) from fuzzywuzzy 0.18.0
(requirements.txt:14).  The package is not in this image and there is no network, so refshim installs this
restatement of its published algorithm under the stub module `fuzzywuzzy.fuzz`; everything around it -- the
reference's rule scoring, tiers and text-similarity ordering -- is the reference's own code.  Written in the
library's decorator-by-decorator form, on difflib.SequenceMatcher (the backend fuzzywuzzy falls back to without
python-Levenshtein, which the reference does not require).  Parity pinned by the package's documented known answers
(tests/test_retrieval_host.py); "unpinned" against the package itself.
"""
import difflib


def _intr(n):
    return int(round(n))


def partial_ratio(s1, s2):
    # check_for_none
    if s1 is None or s2 is None:
        return 0
    # check_for_equivalence
    if s1 == s2:
        return 100
    # check_empty_string
    if len(s1) == 0 or len(s2) == 0:
        return 0
    if len(s1) <= len(s2):
        shorter, longer = s1, s2
    else:
        shorter, longer = s2, s1
    blocks = difflib.SequenceMatcher(None, shorter, longer).get_matching_blocks()
    scores = []
    for block in blocks:
        long_start = block[1] - block[0] if (block[1] - block[0]) > 0 else 0
        long_end = long_start + len(shorter)
        r = difflib.SequenceMatcher(None, shorter, longer[long_start:long_end]).ratio()
        if r > .995:
            return 100
        scores.append(r)
    return _intr(100 * max(scores))
