"""Scorer packaging: KenLM binary + vocabulary + alphabet -> `.scorer` (the reference's `generate_scorer_package`).

Restates native_client/generate_scorer_package.cpp:16-100 and Scorer::{fill_dictionary, save_dictionary}
(ctcdecode/scorer.cpp:398-437, 238-269; decoder_utils.cpp:107-151): every vocabulary word that can be spelled with
the alphabet becomes the label path  word labels + SPACE  (ilabel = label + 1, 0 being OpenFst's epsilon); the set of
paths is turned into the minimal deterministic acceptor and appended to the LM bytes behind the 'TRIE' header
(magic, version 6, is_utf8 byte, alpha, beta as f64) in OpenFst's aligned ConstFst<StdArc> binary layout.

The reference gets the acceptor from RmEpsilon + Determinize + Minimize; the minimal DFA of a finite language is unique
up to state numbering, and the decoder only ever asks "start state", "arc for label" and "is final", so numbering
states breadth-first from the start (arcs sorted by label) yields a package the reference loads and decodes with
identically -- tests/test_scorer_package.py checks exactly that against the compiled reference.

Word mode and bytes-output (UTF-8) mode (generate_scorer_package.cpp:27-50: inferred from the vocabulary -- all words one
code point long -- unless forced; UTF-8 mode spells words in BYTES over the 255-label UTF8Alphabet and appends no space,
decoder_utils.cpp:108-131).  Host-side tool, not on the hot path."""
import struct

KENLM_MAGIC = b"mmap lm http://kheafield.com/code format version 5\n\x00"
TRIE_MAGIC = 0x54524945          # 'TRIE', scorer.cpp:17
TRIE_FILE_VERSION = 6            # scorer.cpp:18
FST_MAGIC = 2125659606           # kFstMagicNumber, fst.h
ALIGN = 16                       # MappedFile::kArchAlignment

# OpenFst property bits (fst/properties.h) of a minimal, epsilon-free, label-sorted, unweighted acyclic acceptor
K_EXPANDED = 0x1
K_ACCEPTOR, K_I_DET, K_O_DET, K_NO_EPS = 0x10000, 0x40000, 0x100000, 0x800000
K_NO_IEPS, K_NO_OEPS, K_ILABEL_SORTED, K_OLABEL_SORTED = 0x2000000, 0x8000000, 0x10000000, 0x40000000
K_UNWEIGHTED, K_ACYCLIC, K_INITIAL_ACYCLIC = 0x200000000, 0x800000000, 0x2000000000
K_TOP_SORTED, K_NOT_TOP_SORTED = 0x4000000000, 0x8000000000
K_ACCESSIBLE, K_COACCESSIBLE, K_STRING, K_NOT_STRING = 0x10000000000, 0x40000000000, 0x100000000000, 0x200000000000
K_UNWEIGHTED_CYCLES = 0x800000000000


def _split_labels(word, label_to_id):
    """split_into_codepoints + char_map lookup (decoder_utils.cpp:113-124); None if a character is not in the alphabet."""
    out = []
    for ch in word:
        i = label_to_id.get(ch)
        if i is None:
            return None
        out.append(i + 1)
    return out


def build_dictionary(words, labels, utf8=False):
    """Minimal DFA over ilabels (= label + 1) accepting {spelling(w) + SPACE} -- or, in UTF-8 mode, {bytes(w)} over the
    UTF8Alphabet (label = byte - 1, so ilabel = the byte).  Returns (start, final flags, arcs) with arcs[state] = sorted
    list of (ilabel, next state), states numbered breadth-first from the start state 0."""
    label_to_id = {l: i for i, l in enumerate(labels)} if not utf8 else {}
    space = None if utf8 else label_to_id[" "] + 1
    # --- trie of the label paths
    children = [{}]
    final = [False]
    n_words = 0
    for w in sorted(set(words)):
        if w in ("<s>", "</s>", "<unk>") or not w:     # scorer.cpp:407
            continue
        path = [b for b in w.encode("utf-8")] if utf8 else _split_labels(w, label_to_id)
        if path is None or (utf8 and 0 in path):
            continue
        n_words += 1
        s = 0
        for lab in (path if utf8 else path + [space]):
            nxt = children[s].get(lab)
            if nxt is None:
                nxt = len(children)
                children.append({})
                final.append(False)
                children[s][lab] = nxt
            s = nxt
        final[s] = True
    # --- minimise bottom-up: states with the same finality and the same (label -> class) map are one class
    order = []
    stack = [0]
    while stack:                       # iterative post-order
        s = stack.pop()
        order.append(s)
        stack.extend(children[s].values())
    cls_of = [None] * len(children)
    registry = {}
    for s in reversed(order):          # children before parents
        sig = (final[s], tuple(sorted((lab, cls_of[c]) for lab, c in children[s].items())))
        cls_of[s] = registry.setdefault(sig, len(registry))
    rep = {}
    for s in order:
        rep.setdefault(cls_of[s], s)
    # --- number the classes breadth-first from the start, arcs in label order
    start_cls = cls_of[0]
    new_id = {start_cls: 0}
    queue = [start_cls]
    arcs, fin = [], []
    qi = 0
    while qi < len(queue):
        c = queue[qi]
        qi += 1
        s = rep[c]
        row = []
        for lab in sorted(children[s]):
            cc = cls_of[children[s][lab]]
            if cc not in new_id:
                new_id[cc] = len(new_id)
                queue.append(cc)
            row.append((lab, new_id[cc]))
        arcs.append(row)
        fin.append(final[s])
    return 0, fin, arcs, n_words


def _const_fst_bytes(start, fin, arcs, offset):
    """ConstFst<StdArc>::Write with FstWriteOptions::align = true (const-fst.h:300-360, fst.cc:84-120): `offset` is the
    position in the file at which the FST starts (alignment is relative to the stream position)."""
    n_states = len(fin)
    n_arcs = sum(len(r) for r in arcs)
    top_sorted = all(nxt > s for s, row in enumerate(arcs) for _, nxt in row)
    is_string = all(len(r) <= 1 for r in arcs)
    props = (K_EXPANDED | K_ACCEPTOR | K_I_DET | K_O_DET | K_NO_EPS | K_NO_IEPS | K_NO_OEPS | K_ILABEL_SORTED |
             K_OLABEL_SORTED | K_UNWEIGHTED | K_ACYCLIC | K_INITIAL_ACYCLIC | K_ACCESSIBLE | K_COACCESSIBLE |
             K_UNWEIGHTED_CYCLES | (K_TOP_SORTED if top_sorted else K_NOT_TOP_SORTED) | (K_STRING if is_string else K_NOT_STRING))
    out = bytearray()
    out += struct.pack("<i", FST_MAGIC)
    for s in (b"const", b"standard"):
        out += struct.pack("<i", len(s)) + s
    out += struct.pack("<iiQqqq", 1, 4, props, start, n_states, n_arcs)   # kAlignedFileVersion, IS_ALIGNED

    def align():
        pad = (-(offset + len(out))) % ALIGN
        out.extend(b"\x00" * pad)
    align()
    pos = 0
    for s in range(n_states):
        w = 0.0 if fin[s] else float("inf")            # TropicalWeight::One() / Zero()
        out += struct.pack("<fIIII", w, pos, len(arcs[s]), 0, 0)
        pos += len(arcs[s])
    align()
    for row in arcs:
        for lab, nxt in row:
            out += struct.pack("<iifi", lab, lab, 0.0, nxt)
    return bytes(out)


def looks_char_based(vocab_words):
    """generate_scorer_package.cpp:29-40: every vocabulary word is one code point long."""
    return all(len(w) <= 1 for w in vocab_words)


def create_scorer_package(lm_path, vocab_words, labels, package_path, default_alpha, default_beta, utf8=None):
    """generate_scorer_package --lm LM --vocab VOCAB --package OUT --default_alpha A --default_beta B
    [--force_bytes_output_mode].  `labels`: the model's alphabet in label order (word mode: must contain the space; UTF-8
    mode: ignored, the alphabet is the 255 byte values).  `utf8` None = inferred from the vocabulary like the reference.
    Returns (#words in the dictionary, #states, #arcs).  The LM must be a KenLM binary built with -v (no vocabulary strings
    behind the search section)."""
    if utf8 is None:
        utf8 = looks_char_based(vocab_words)
    lm = open(lm_path, "rb").read()
    if not lm.startswith(KENLM_MAGIC):
        raise ValueError("not a KenLM binary (format version 5): %s" % lm_path)
    if not utf8 and (labels is None or " " not in labels):
        raise ValueError("word-mode scorers need an alphabet with a space label")
    if struct.pack("<i", TRIE_MAGIC) in lm[-(1 << 16):] and lm.rfind(struct.pack("<ii", TRIE_MAGIC, TRIE_FILE_VERSION)) >= 0:
        raise ValueError("the LM file already carries a 'TRIE' dictionary section: pass the bare KenLM binary")
    start, fin, arcs, n_words = build_dictionary(vocab_words, labels, utf8)
    head = struct.pack("<iiBdd", TRIE_MAGIC, TRIE_FILE_VERSION, 1 if utf8 else 0, float(default_alpha), float(default_beta))
    fst = _const_fst_bytes(start, fin, arcs, len(lm) + len(head))
    with open(package_path, "wb") as f:
        f.write(lm)
        f.write(head)
        f.write(fst)
    return n_words, len(fin), sum(len(r) for r in arcs)


def read_alphabet(path):
    """Alphabet::init (alphabet.cc:42-68): one label per line; a line that is exactly `\\#` is the label '#', other
    lines starting with '#' are comments, empty lines are skipped."""
    labels = []
    for line in open(path, encoding="utf-8"):
        line = line.rstrip("\r\n")
        if line == "\\#":
            line = "#"
        elif line.startswith("#"):
            continue
        if line == "":
            continue
        labels.append(line)
    return labels


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="Create a .scorer package from a KenLM binary and a vocabulary.")
    ap.add_argument("--alphabet", help="alphabet.txt: one label per line, '#' comments (alphabet.cc:42-68); word mode only")
    ap.add_argument("--force_bytes_output_mode", type=int, choices=(0, 1), default=None,
                    help="1 = UTF-8 bytes mode, 0 = word mode; default: inferred from the vocabulary like the reference")
    ap.add_argument("--lm", required=True)
    ap.add_argument("--vocab", required=True)
    ap.add_argument("--package", required=True)
    ap.add_argument("--default_alpha", type=float, required=True)
    ap.add_argument("--default_beta", type=float, required=True)
    a = ap.parse_args(argv)
    words = open(a.vocab, encoding="utf-8").read().split()
    utf8 = looks_char_based(words) if a.force_bytes_output_mode is None else bool(a.force_bytes_output_mode)
    if not utf8 and not a.alphabet:
        ap.error("word mode needs --alphabet (the reference asks for --checkpoint with an alphabet.txt)")
    labels = read_alphabet(a.alphabet) if a.alphabet else None
    n, ns, na = create_scorer_package(a.lm, words, labels, a.package, a.default_alpha, a.default_beta, utf8)
    print("%d words, %d states, %d arcs -> %s" % (n, ns, na, a.package))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
