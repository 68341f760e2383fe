"""One-off source transformation: <<<>>> launches -> launch_pdl(), EFFDET_PDL_SYNC() at kernel entry."""
import re, sys

def convert(path, only=None, skip=()):
    s = open(path).read()
    # ---- launches
    out, i, converted = [], 0, set()
    while True:
        j = s.find("<<<", i)
        if j < 0:
            out.append(s[i:]); break
        # kernel expression: scan back over identifier / template args
        k = j
        depth = 0
        while k > 0:
            c = s[k - 1]
            if c == '>': depth += 1
            elif c == '<': depth -= 1
            elif depth == 0 and not (c.isalnum() or c in "_:"): break
            k -= 1
        kern = s[k:j]
        name = re.match(r"[A-Za-z_0-9:]+", kern).group(0)
        e = s.find(">>>", j)
        cfg = s[j + 3:e]
        # args: balanced parens after >>>
        assert s[e + 3] == '(', (path, s[e:e + 40])
        d, m = 0, e + 3
        while True:
            if s[m] == '(': d += 1
            elif s[m] == ')':
                d -= 1
                if d == 0: break
            m += 1
        args = s[e + 4:m]
        if (only and name not in only) or name in skip:
            out.append(s[i:m + 1]); i = m + 1; continue
        parts, d2, cur = [], 0, ""
        for ch in cfg:
            if ch in "(<[": d2 += 1
            if ch in ")>]": d2 -= 1
            if ch == ',' and d2 == 0: parts.append(cur.strip()); cur = ""
            else: cur += ch
        parts.append(cur.strip())
        while len(parts) < 4: parts.append("0")
        g, b, sm, st = parts
        call = "launch_pdl(%s, dim3(%s), dim3(%s), %s, %s, %s)" % (kern, g, b, sm, st, args)
        # statement or expression context?
        rest = s[m + 1:m + 2]
        prev = s[:k].rstrip()[-1:]
        if rest == ';' and prev not in "(,":
            out.append(s[i:k] + "EFFDET_CUDA(" + call + ")")
        else:       # expression context (inside a dispatch macro): EFFDET_LAUNCHED() after it reports failures
            out.append(s[i:k] + "(void)" + call)
        converted.add(name)
        i = m + 1
    s = "".join(out)
    # ---- kernel entries
    for name in sorted(converted):
        pat = re.compile(r"(__global__[^;{]*?\b%s\s*\([^{;]*?\)\s*\{)" % re.escape(name), re.S)
        ms = list(pat.finditer(s))
        assert ms, (path, name)
        for mm in reversed(ms):
            s = s[:mm.end()] + "\n    EFFDET_PDL_SYNC();" + s[mm.end():]
    open(path, "w").write(s)
    return converted

if __name__ == "__main__":
    print(convert(sys.argv[1], only=set(sys.argv[2].split(",")) if len(sys.argv) > 2 and sys.argv[2] else None))
