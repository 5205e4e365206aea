"""Index of the C ABI: every function include/fluidgym_b200.h declares, with the comment that precedes it (which names the
reference interface it replaces, file:line).  ``python tools/abi_index.py`` prints the markdown table kept as the appendix
of INTEGRATION.md; tests/test_abi.py checks that the appendix names every declared function."""
import os
import re

HEADER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "fluidgym_b200.h")


def declarations(path=HEADER):
    """[(name, comment)] in header order; comment = the block comment directly above the declaration, or the one of the
    declaration directly above it when they share a comment (e.g. create / destroy pairs, the _scalar variants)."""
    src = open(path).read()
    token = re.compile(r"/\*(.*?)\*/|^[A-Za-z_][^;{}#/]*?\b(fgb_[a-z0-9_]+)\s*\((?:[^;{}/]|/\*.*?\*/)*\)\s*;", re.S | re.M)
    out, prev_end, prev_comment = [], 0, ""
    for m in token.finditer(src):
        gap = src[prev_end:m.start()]
        adjacent = not gap.strip() and gap.count("\n") <= 1
        if m.group(2) is None:
            text = re.sub(r"\s*\n\s*\*\s?", " ", m.group(1)).strip()
            own_line = src.rfind("\n", 0, m.start()) + 1 == m.start()        # a comment that starts its line, not a trailing remark
            prev_comment = text if own_line else ""
        else:
            if not adjacent:
                prev_comment = ""
            out.append((m.group(2), prev_comment))
        prev_end = m.end()
    return out


def first_sentence(text, limit=230):
    text = re.sub(r"\s+", " ", text)
    text = re.sub(r"^-+\s*", "", text)
    text = re.sub(r"\s*-{4,}\s*", ". ", text).strip()                   # section rules: "---- title ------ body"
    if len(text) <= limit:
        return text
    cut = text.rfind(". ", 0, limit)
    return text[:cut + 1] if cut > 60 else text[:limit].rsplit(" ", 1)[0] + " ..."


def table():
    rows = ["| entry point | replaces (reference interface, file:line) |", "|---|---|"]
    for name, comment in declarations():
        rows.append("| `%s` | %s |" % (name, first_sentence(comment).replace("|", "/")))
    return "\n".join(rows)


if __name__ == "__main__":
    print(table())
