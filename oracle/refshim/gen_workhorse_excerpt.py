#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.  Cuts the step between the phases out of the reference's OWN src/crass/WorkHorse.cpp, verbatim,
so that the checker compiles and runs the reference's code for it instead of a restatement (WorkHorse.cpp as a whole needs
the Xerces headers, which this image does not have; these functions need none of them):

    sortLengthAssending, includeSubstring, isNotEmpty          WorkHorse.cpp:72-92
    WorkHorse::removeRedundantRepeats                          WorkHorse.cpp:612-645
    WorkHorse::createNonRedundantSet                           WorkHorse.cpp:648-709
    WorkHorse::clusterDRReads                                  WorkHorse.cpp:1404-1637

    gen_workhorse_excerpt.py <reference>/src/crass/WorkHorse.cpp <out.inc>

The output goes to oracle/_ref/ (git-ignored): no reference source is copied into the repository.  Functions are found by
their signature and cut by brace matching, so the script fails loudly if the reference changes shape."""
import re
import sys

WANT = [r"^bool\s+sortLengthAssending\s*\(", r"^bool\s+includeSubstring\s*\(", r"^bool\s+isNotEmpty\s*\(",
        r"^void\s+WorkHorse::removeRedundantRepeats\s*\(", r"^Vecstr\s*\*\s*WorkHorse::createNonRedundantSet\s*\(",
        r"^bool\s+WorkHorse::clusterDRReads\s*\("]


def cut(lines, start):
    depth, seen, out = 0, False, []
    for i in range(start, len(lines)):
        line = lines[i]
        out.append(line)
        code = re.sub(r'"(\\.|[^"\\])*"', '""', line)          # braces inside string literals do not count
        code = re.sub(r"'(\\.|[^'\\])'", "''", code)
        code = code.split("//")[0]
        depth += code.count("{") - code.count("}")
        seen = seen or "{" in code
        if seen and depth == 0:
            return out, i + 1
    raise SystemExit("unbalanced braces after line %d" % (start + 1))


def main():
    src, dst = sys.argv[1], sys.argv[2]
    lines = open(src, encoding="latin-1").read().split("\n")
    parts = ["// GENERATED from %s by oracle/refshim/gen_workhorse_excerpt.py -- the reference's own code, not part of the repository\n" % src]
    for pat in WANT:
        hits = [i for i, l in enumerate(lines) if re.match(pat, l)]
        if len(hits) != 1:
            raise SystemExit("expected exactly one definition matching %r, found %d" % (pat, len(hits)))
        body, end = cut(lines, hits[0])
        parts.append("// ---- %s:%d-%d\n%s\n" % (src.split("/")[-1], hits[0] + 1, end, "\n".join(body)))
    open(dst, "w", encoding="latin-1").write("\n".join(parts))


if __name__ == "__main__":
    main()
