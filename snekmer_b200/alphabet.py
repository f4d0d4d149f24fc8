"""alphabet: amino-acid reduction (AAR) alphabets — drop-in for ``snekmer.alphabet``.

Same public names, values and error behaviour as the reference module
(snekmer/alphabet.py:13-266): ``ALPHABETS``, ``FULL_ALPHABETS``,
``ALPHABET_ORDER``, ``ALPHABET_ID``, ``ALPHABET2ID``, ``check_valid``,
``get_alphabet``, ``get_alphabet_name``, ``get_alphabet_keys``,
``get_alphabets``.  Additions for the device path: ``register_alphabet``,
``symbols`` (canonical sorted output symbols = digit order of the integer k-mer
codes) and ``lut`` / ``charmap`` (256-entry byte tables fed to the kernels).
"""
from __future__ import annotations

from typing import Dict, Set, Union

StandardAlphabet = "AILMVFYWSTQNCHDEKRGP"
AA_SELF_MAPPING = {a: a for a in StandardAlphabet}
PTM_CHARS = "-_!^#$@.%&"
PTM_SELF_MAPPING = {c: c for c in PTM_CHARS}

ALPHABET_ORDER = dict(enumerate(("hydro", "standard", "solvacc", "hydrocharge", "hydrostruct", "miqs")))


def _spec(text: str, keys: str) -> Dict[str, str]:
    """'RESIDUES>S ...' → {'RESIDUES': 'S', ..., '_keys': keys} (reference layout)."""
    d = {}
    for item in text.split():
        src, dst = item.split(">")
        d[src] = dst
    d["_keys"] = keys
    return d


# Group strings and symbols are the reference's (alphabet.py:31-85); they are data,
# including the two quirks that parity depends on: "hydrocharge" has no E and
# lists N twice (the later group wins), "hydrostruct" emits the symbol B.
ALPHABETS: Dict[str, Dict[str, str]] = {
    "hydro": _spec("SFTNKYEQCWPHDR>S VMLAIG>V", "SV"),
    "standard": _spec("AGILMV>A PH>P FWY>F NQST>N DE>D KR>K C>C", "APFNDKC"),
    "solvacc": _spec("CILMVFWY>C AGHST>A PDEKNQR>P", "CAP"),
    "hydrocharge": _spec("SFTNYQCWPH>L VMLAIG>H KNDR>C", "LHC"),
    "hydrostruct": _spec("SFTNKYEQCWHDR>L VMLAI>H PG>B", "LHB"),
    "miqs": _spec("A>A C>C DEN>D FWY>F G>G H>H ILMQV>I KR>K P>P ST>S", "ACDFGHIKPS"),
    "ptm": {**AA_SELF_MAPPING, **PTM_SELF_MAPPING, "_keys": StandardAlphabet + PTM_CHARS},
    "None": AA_SELF_MAPPING,
}


def _long_form(mapping: Dict[str, str]) -> Dict[str, str]:
    out: Dict[str, str] = {}
    for group, symbol in mapping.items():
        if group != "_keys":
            out.update({residue: symbol for residue in group})
    return out


FULL_ALPHABETS: Dict[str, dict] = {name: _long_form(m) for name, m in ALPHABETS.items()}

ALPHABET_ID = {
    f"RED{n}": {v: k for k, v in ALPHABETS[ALPHABET_ORDER[n]].items()} for n in range(len(ALPHABET_ORDER))
}
ALPHABET2ID = {ALPHABET_ORDER[n]: f"RED{n}" for n in range(len(ALPHABET_ORDER))}


def get_alphabets():
    """All alphabet mappings, ``{'alphabet': {'residues': 'symbol', ...}}``."""
    return ALPHABETS


def check_valid(alphabet: Union[str, int]) -> None:
    """Raise ValueError unless `alphabet` names a defined alphabet (name, index or None).

    Like the reference (alphabet.py:121-155) the integer test is
    ``alphabet in range(len(ALPHABETS))``, so 6 and 7 pass here and fail later
    with KeyError in ALPHABET_ORDER."""
    known = (alphabet in range(len(ALPHABETS))) or (alphabet in ALPHABETS)
    if not known and str(alphabet) != "None":
        raise ValueError(
            "Invalid alphabet specified; alphabet must be a string (see snekmer.alphabet) or integer n between"
            f" {min(ALPHABET_ORDER)} and {max(ALPHABET_ORDER)}."
        )


def get_alphabet_name(alphabet: Union[str, int], mapping: dict = ALPHABETS) -> str:
    """Alphabet name for a name / index / None input."""
    check_valid(alphabet)
    if alphabet is None:
        return "None"
    if isinstance(alphabet, int):
        return ALPHABET_ORDER[alphabet]
    return alphabet


def get_alphabet(alphabet: Union[str, int], mapping: dict = ALPHABETS) -> Dict[str, str]:
    """The residue→symbol map of `alphabet` looked up in `mapping`."""
    return mapping[get_alphabet_name(alphabet)]


def get_alphabet_keys(alphabet: Union[str, int], mapping: Dict[str, dict] = FULL_ALPHABETS) -> Set[str]:
    """Set of output symbols of `alphabet` (drops a '_keys' entry in place, as the reference does)."""
    alphabet_map = get_alphabet(alphabet, mapping)
    if "_keys" in alphabet_map:
        alphabet_map.pop("_keys")
    return set(alphabet_map.values())


# ---------------------------------------------------------------------------
# additions for the device path
# ---------------------------------------------------------------------------
def register_alphabet(name: str, groups: Dict[str, str]) -> None:
    """Add a custom alphabet, e.g. ``register_alphabet("syn6", {"AGILMV": "A", ...})``.

    After registration ``check_valid(name)`` accepts it by name, exactly as if it
    had been part of ALPHABETS."""
    if not isinstance(name, str) or not name:
        raise ValueError("alphabet name must be a non-empty string")
    groups = {k: v for k, v in groups.items() if k != "_keys"}
    for src, dst in groups.items():
        if len(dst) != 1 or not src:
            raise ValueError("each group must map >=1 residues to ONE output symbol")
    keys = "".join(dict.fromkeys(groups.values()))
    ALPHABETS[name] = {**groups, "_keys": keys}
    FULL_ALPHABETS[name] = _long_form(groups)


def symbols(alphabet: Union[str, int]) -> str:
    """Output symbols in canonical (sorted) order: digit i of a k-mer code is symbols[i]."""
    return "".join(sorted(set(FULL_ALPHABETS[get_alphabet_name(alphabet)].values())))


def residue_map(alphabet: Union[str, int]) -> Dict[str, str]:
    """Long-form residue→symbol dict without '_keys'."""
    return {k: v for k, v in FULL_ALPHABETS[get_alphabet_name(alphabet)].items() if k != "_keys"}


def lut(alphabet: Union[str, int]) -> bytes:
    """256-byte residue→symbol-index table (0xFF = invalid), built by the C library."""
    from . import _native

    m = residue_map(alphabet)
    src = "".join(m.keys())
    dst = "".join(m.values())
    return _native.lut_build(src, dst, symbols(alphabet))


def charmap(alphabet: Union[str, int]) -> bytes:
    """256-byte residue→reduced-character table (``str.translate`` semantics: unmapped bytes stay)."""
    table = bytearray(range(256))
    for src, dst in residue_map(alphabet).items():
        table[ord(src)] = ord(dst)
    return bytes(table)
