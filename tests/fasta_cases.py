"""Seeded FASTA texts for the ingest path (parse_fasta.cpp:9-89 + Sequence::operator=(deque<char>) + Sequence::defline), shared by
the golden generator and the tests.  Every quirk of the reference's reader that changes the result is in here: gzgets chunks of
2047 bytes, a chunk holding '>' anywhere is a defline, CR LF, blank lines, white space inside lines, lower case, IUPAC, '-' (EOS),
'[w=...]' weights, the length window, the ignore list, a missing final newline, a last record that is empty."""
import numpy as np

LETTERS = "ACGTacgtUuMRSVWYHKDBNmrsvwyhkdbnIiXx-"


def _seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(list(alphabet), size=n))


def _wrap(s, width, eol="\n"):
    return eol.join(s[i:i + width] for i in range(0, len(s), width)) + eol


class FastaCase:
    def __init__(self, name, files, min_len=0, max_len=1 << 40, ignore=()):
        self.name, self.files, self.min_len, self.max_len, self.ignore = name, [f.encode() for f in files], min_len, max_len, list(ignore)


def cases():
    rng = np.random.default_rng(77)
    out = []
    # plain: several records, line width 70, one file
    t = "".join(">t%d some text\n%s" % (i, _wrap(_seq(rng, int(rng.integers(200, 3000))), 70)) for i in range(12))
    out.append(FastaCase("plain", [t]))
    # two files, CR LF, blank lines, spaces and tabs inside lines, lower case + IUPAC + U + I/X + '-', weights, no final newline
    recs = []
    for i in range(9):
        body = _seq(rng, int(rng.integers(50, 900)), LETTERS if i % 2 else "ACGTacgtN")
        lines = _wrap(body, int(rng.integers(7, 120)), "\r\n" if i % 3 == 0 else "\n")
        lines = lines.replace("A", "A ", 2).replace("C", "\tC", 1)
        if i % 4 == 1:
            lines = "\n\n" + lines + "  \n"
        w = ["", " [w=2.5]", " [ w = 0.25 ] tail", " [w=3]", " [[w=1e1]", " [w=.5 ]", " [w= x]", " [W=7.75] [w=9]", " [w=2"][i]
        recs.append(">rec%d%s%s%s" % (i, w, "\r\n" if i % 3 == 0 else "\n", lines))
    f1, f2 = "".join(recs[:5]), "".join(recs[5:])
    out.append(FastaCase("messy_two_files", [f1, f2.rstrip("\n")]))
    # length window + ignore list + consecutive deflines + an empty last record
    t = ""
    for i, n in enumerate([10, 400, 0, 120, 2500, 60]):
        t += ">s%d %s\n" % (i, "Plasmid pX" if i == 3 else "chromosome")
        if n:
            t += _wrap(_seq(rng, n), 60)
    t += ">last one has no sequence\n"
    out.append(FastaCase("window_ignore", [t], min_len=50, max_len=2000, ignore=["plasmid"]))
    # gzgets semantics: lines longer than 2047 bytes come in chunks; '>' inside a sequence line turns that CHUNK into a defline;
    # a very long sequence line; a file that does not start with a defline
    long_line = _seq(rng, 7000)
    t = _wrap(_seq(rng, 100), 50) + ">a\n" + long_line + "\n>b\n" + _seq(rng, 300) + "\n" + _seq(rng, 40) + ">" + _seq(rng, 20) + "\n" + _seq(rng, 90) + "\n"
    out.append(FastaCase("chunks", [t]))
    # '>' in the second gzgets chunk of a 5000-byte line: bytes 2047..4093 of that line are a "defline", the rest is sequence;
    # consecutive deflines; an empty record after the last defline, kept because nothing asks for a minimum length
    line = _seq(rng, 3000) + ">" + _seq(rng, 1999)
    t = ">x\n" + _seq(rng, 64) + "\n" + line + "\n" + _seq(rng, 33) + "\n>y\n>z [w=4]\n" + _seq(rng, 77) + "\n>empty\n"
    out.append(FastaCase("chunk_defline", [t]))
    # odd lengths and very short records (pad nibble, one byte)
    t = "".join(">o%d\n%s\n" % (i, _seq(rng, n)) for i, n in enumerate([1, 2, 3, 31, 32, 33, 63, 65, 4097, 8191]))
    out.append(FastaCase("odd", [t]))
    return out


class GroupCase:
    """append_fasta_group (parse_fasta.cpp:91-169): files, the group of each file, the record window, the pad between records"""
    def __init__(self, name, files, file_group, min_len=0, max_len=1 << 40, num_pad=1, ignore=()):
        self.name, self.files, self.file_group = name, [f.encode() for f in files], list(file_group)
        self.min_len, self.max_len, self.num_pad, self.ignore = min_len, max_len, num_pad, list(ignore)


def group_cases():
    rng = np.random.default_rng(177)
    out = []

    def fa(names_lens, alphabet="ACGT", width=70):
        return "".join(">%s\n%s" % (nm, _wrap(_seq(rng, n, alphabet), width) if n else "") for nm, n in names_lens)

    # three groups: contigs of a draft genome in one file, a genome split over two files, a single-record group; odd and even
    # record lengths (a pad lands on either nibble of a byte), records across file boundaries
    f0 = fa([("c1", 301), ("c2", 120), ("c3", 77)])
    f1 = fa([("a1", 250)])
    f2 = fa([("a2", 33), ("a3", 500)])
    f3 = fa([("solo", 999)])
    out.append(GroupCase("three_groups", [f0, f1, f2, f3], [0, 1, 1, 2]))
    # length window, ignore list, a group that keeps nothing (dropped), IUPAC / lower case / '-' inside records, two pads, CR LF
    g0 = fa([("k1 plasmid", 400), ("k2", 30), ("k3", 200)], "ACGTacgtNRY-")
    g1 = fa([("short1", 10), ("short2", 20)])
    g2 = fa([("m1", 150), ("m2 PLASMID x", 300), ("m3", 151)], width=61).replace("\n", "\r\n")
    g3 = fa([("z", 90)])
    out.append(GroupCase("window_ignore_pad2", [g0, g1, g2, g3], [0, 1, 2, 2], min_len=50, max_len=380, num_pad=2, ignore=["plasmid"]))
    # gzgets chunking inside a group: a 5000-base line, a '>' in the middle of a sequence line, no final newline, a file that does
    # not start with a defline, an empty last record
    t = _seq(rng, 64) + "\n>p\n" + _seq(rng, 5000) + "\n>q\n" + _seq(rng, 40) + ">" + _seq(rng, 20) + "\n" + _seq(rng, 95)
    u = ">r\n" + _wrap(_seq(rng, 333), 50) + ">empty\n"
    out.append(GroupCase("chunks", [t, u], [0, 0]))
    return out
