"""Minimal stand-in for the third-party `inflection` package (not installed in this image, no network).

scripts/main.py:19,109-115 only calls `inflection.pluralize(key)` on the dataset's dict keys.  This module implements
English pluralisation by the usual ordered suffix rules (uncountables, irregulars, then regex rules tried last to
first) — enough for any identifier-like key.  tools/run_main.py puts this directory on sys.path ONLY when the real
package cannot be imported.
"""
import re

_UNCOUNTABLE = {"equipment", "information", "rice", "money", "species", "series", "fish", "sheep", "jeans", "police"}
_IRREGULAR = {"person": "people", "man": "men", "child": "children", "sex": "sexes", "move": "moves", "cow": "kine",
              "zombie": "zombies"}
_RULES = [      # (pattern, replacement), most specific first
    (r"(?i)(quiz)$", r"\1zes"),
    (r"(?i)^(oxen)$", r"\1"),
    (r"(?i)^(ox)$", r"\1en"),
    (r"(?i)(m|l)ice$", r"\1ice"),
    (r"(?i)(m|l)ouse$", r"\1ice"),
    (r"(?i)(passer)s?by$", r"\1sby"),
    (r"(?i)(matr|vert|ind)(?:ix|ex)$", r"\1ices"),
    (r"(?i)(x|ch|ss|sh)$", r"\1es"),
    (r"(?i)([^aeiouy]|qu)y$", r"\1ies"),
    (r"(?i)(hive)$", r"\1s"),
    (r"(?i)([lr])f$", r"\1ves"),
    (r"(?i)([^f])fe$", r"\1ves"),
    (r"(?i)sis$", "ses"),
    (r"(?i)([ti])a$", r"\1a"),
    (r"(?i)([ti])um$", r"\1a"),
    (r"(?i)(buffal|potat|tomat)o$", r"\1oes"),
    (r"(?i)(bu)s$", r"\1ses"),
    (r"(?i)(alias|status)$", r"\1es"),
    (r"(?i)(octop|vir)i$", r"\1i"),
    (r"(?i)(octop|vir)us$", r"\1i"),
    (r"(?i)^(ax|test)is$", r"\1es"),
    (r"(?i)s$", "s"),
    (r"$", "s"),
]


def pluralize(word: str) -> str:
    if not word or word.lower() in _UNCOUNTABLE or any(re.search(rf"(?i)\b{u}\Z", word) for u in _UNCOUNTABLE):
        return word
    for singular, plural in _IRREGULAR.items():
        if re.search(rf"(?i)({singular[0]}){singular[1:]}$", word):
            return re.sub(rf"(?i)({singular[0]}){singular[1:]}$", r"\1" + plural[1:], word)
        if re.search(rf"(?i)({plural[0]}){plural[1:]}$", word):
            return word
    for pattern, replacement in _RULES:
        if re.search(pattern, word):
            return re.sub(pattern, replacement, word)
    return word
