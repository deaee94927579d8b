"""One-off source transform: rewrite `kernel<T...><<<grid, block, smem, stream>>>(args)` launch sites of the named kernels
into `mdm_launch(kernel<T...>, grid, block, smem, stream, args)` (csrc/common.cuh: the programmatic-dependent-launch
attribute), refusing kernels whose body does not call pdl_enter() / pdl_wait().  usage: pdl_convert.py file.cu name..."""
import re, sys

def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[{": depth += 1
        if ch in ")>]}": depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out

def convert(path, names):
    s = open(path).read()
    for n in names:
        m = re.search(r"__global__[^;{]*?\b%s\s*\(" % n, s, re.S)
        assert m, "kernel %s not found in %s" % (n, path)
        # body: from the first '{' after the signature to its matching '}'
        i = s.index("{", s.index(")", m.end()))
        # the signature may contain parentheses inside (none of ours do after the parameter list)
        depth, j = 0, i
        while True:
            if s[j] == "{": depth += 1
            if s[j] == "}":
                depth -= 1
                if depth == 0: break
            j += 1
        assert re.search(r"pdl_(enter|wait)\(\)", s[i:j]), "kernel %s has no pdl_enter()" % n
    pos, count = 0, 0
    while True:
        k = s.find("<<<", pos)
        if k < 0: break
        # kernel expression: walk back over optional template args and identifier
        b = k
        if s[b - 1] == ">":
            depth = 0
            while True:
                b -= 1
                if s[b] == ">": depth += 1
                if s[b] == "<":
                    depth -= 1
                    if depth == 0: break
        e = b
        while s[e - 1].isalnum() or s[e - 1] == "_": e -= 1
        kexpr = s[e:k]
        kname = re.match(r"\w+", kexpr).group(0)
        c_end = s.index(">>>", k)
        # the config itself may contain '>' of casts, find the >>> followed by '('
        while s[c_end + 3] != "(":
            c_end = s.index(">>>", c_end + 1)
        cfg = split_top(s[k + 3:c_end])
        a0 = c_end + 3
        depth, a1 = 0, a0
        while True:
            if s[a1] == "(": depth += 1
            if s[a1] == ")":
                depth -= 1
                if depth == 0: break
            a1 += 1
        if kname not in names:
            pos = a1
            continue
        while len(cfg) < 4: cfg.append("0")
        args = s[a0 + 1:a1].strip()
        new = "mdm_launch(%s, %s, %s, %s, %s%s)" % (kexpr, cfg[0], cfg[1], cfg[2], cfg[3], (", " + args) if args else "")
        s = s[:e] + new + s[a1 + 1:]
        pos = e + len(new)
        count += 1
    open(path, "w").write(s)
    print(path, "converted", count, "launch sites")

if __name__ == "__main__":
    convert(sys.argv[1], sys.argv[2:])
