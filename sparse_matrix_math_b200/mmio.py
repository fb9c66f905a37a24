"""Matrix Market loader with the semantics of SMM::loadMatrixMarketMatrix (H:2531-2609): host-side setup code.

coordinate x {real, integer} x symmetric only; banner token case-sensitive, the others lower-cased; off-diagonal
entries mirrored; explicit zeros kept; values parsed straight to float32 (correctly rounded, like operator>>).
"""
import numpy as np


def load_matrix_market(path, out, extended=False):
    """extended=False: the reference's loader.  extended=True (extension): also `general` /
    `skew-symmetric` structure and the `pattern` field (every stored entry is 1)."""
    from .binding import MatrixLoadStatus as S
    try:
        f = open(path, "r")
    except OSError:
        return S.FAILED_TO_OPEN_FILE
    with f:
        text = f.read()
    pos = 0

    def token():
        nonlocal pos
        n = len(text)
        while pos < n and text[pos].isspace():
            pos += 1
        if pos >= n:
            return None
        b = pos
        while pos < n and not text[pos].isspace():
            pos += 1
        return text[b:pos]

    if token() != "%%MatrixMarket":
        return S.PARSE_ERROR_MMX_FILE_MISSING_BANNER
    if (token() or "").lower() != "matrix":
        return S.PARSE_ERROR_MMX_FILE_UNSUPPORTED_TYPE
    if (token() or "").lower() != "coordinate":
        return S.PARSE_ERROR_MMX_FILE_UNSUPPORTED_FORMAT
    el_type = (token() or "").lower()
    pattern = extended and el_type == "pattern"
    if el_type not in ("real", "integer") and not pattern:
        return S.PARSE_ERROR_MMX_FILE_UNSUPPORTED_EL_TYPE
    structure = (token() or "").lower()
    general = extended and structure == "general"
    skew = extended and structure == "skew-symmetric"
    if structure != "symmetric" and not general and not skew:
        return S.PARSE_ERROR_MMX_FILE_UNSUPPORTED_STRUCTURE
    # H:2576-2578: drop every line that starts with '%' or whitespace
    n = len(text)
    while pos < n and (text[pos] == "%" or text[pos].isspace()):
        nl = text.find("\n", pos)
        pos = n if nl < 0 else nl + 1
    try:
        rows, cols, _nnz = int(token()), int(token()), int(token())
    except (TypeError, ValueError):
        return S.FAILED_TO_PARSE_FILE
    out.init(rows, cols, _nnz)
    if extended and _nnz == 0:
        return S.SUCCESS
    while True:                                   # H:2588: the body runs at least once
        try:
            r, c = int(token()), int(token())
            v = np.float32(1.0) if pattern else np.float32(token())
        except (TypeError, ValueError):
            return S.FAILED_TO_PARSE_FILE
        r -= 1
        c -= 1
        out.addEntry(r, c, v)
        if not general and r != c:
            out.addEntry(c, r, -v if skew else v)
        while pos < n and text[pos].isspace():    # H:2603-2605
            nl = text.find("\n", pos)
            pos = n if nl < 0 else nl + 1
        if pos >= n:
            break
    return S.SUCCESS
