"""Minimal stand-in for the third-party `newick` package (absent in this image).

Only used by tests/golden/make_golden.py so that the REFERENCE can be imported
(`/root/reference/prosstt/tree.py:14` does `import newick`).  It provides what
the reference touches: `loads(str)` -> [Node], `Node.walk()` (pre-order),
`.name`, `.length` (float, 0.0 when absent), `.descendants`, `.ancestor`
(`/root/reference/prosstt/tree_utils.py:40-55`).
"""


class Node:
    def __init__(self, name=None, length=0.0):
        self.name = name
        self.length = length
        self.descendants = []
        self.ancestor = None

    def walk(self):
        yield self
        for child in self.descendants:
            yield from child.walk()


def _parse(s, pos):
    node = Node()
    if s[pos] == "(":
        pos += 1
        while True:
            child, pos = _parse(s, pos)
            child.ancestor = node
            node.descendants.append(child)
            if s[pos] == ",":
                pos += 1
                continue
            if s[pos] == ")":
                pos += 1
                break
            raise ValueError("bad newick at %d" % pos)
    start = pos
    while pos < len(s) and s[pos] not in ",():;":
        pos += 1
    label = s[start:pos].strip()
    node.name = label if label else None
    if pos < len(s) and s[pos] == ":":
        pos += 1
        start = pos
        while pos < len(s) and s[pos] not in ",();":
            pos += 1
        node.length = float(s[start:pos])
    return node, pos


def loads(text):
    out = []
    for chunk in text.strip().split(";"):
        chunk = chunk.strip()
        if chunk:
            node, _ = _parse(chunk + ";", 0)
            out.append(node)
    return out
