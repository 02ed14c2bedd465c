"""Minimal rooted-tree reader/writer for the one thing tapir needs from a tree: its depth, its total
length and its leaf count, plus branch-length rescaling (tapir/compute.py:59-74, which uses DendroPy).

DendroPy semantics reproduced:
  * ``seed_node.distance_from_tip()``  = max over children of (child's distance + child's edge length),
    a missing length counting as 0;
  * ``tree.length()``                  = sum of all edge lengths that are not None (root edge included);
  * ``leaf_nodes()``                   = nodes without children;
  * rescaling leaves edges with a falsy length (None or 0) untouched (compute.py:70).
Newick and NEXUS (first tree of the first TREES block, TRANSLATE honoured) are supported, as the
``--tree-format`` option of bin/tapir_compute.py:34 allows.
"""
import re


class Node:
    __slots__ = ("label", "length", "children")

    def __init__(self, label=None, length=None):
        self.label = label
        self.length = length
        self.children = []

    def is_leaf(self):
        return not self.children


class Tree:
    def __init__(self, root):
        self.root = root

    def nodes(self):
        stack, out = [self.root], []
        while stack:
            n = stack.pop()
            out.append(n)
            stack.extend(reversed(n.children))
        return out

    def leaf_nodes(self):
        return [n for n in self.nodes() if n.is_leaf()]

    def length(self):
        return sum(n.length for n in self.nodes() if n.length is not None)

    def depth(self):
        """Distance from the root to its farthest tip (iterative post-order; trees can be deep)."""
        dist = {}
        for n in reversed(self.nodes()):          # children are listed after their parent
            dist[id(n)] = max((dist[id(c)] + (c.length or 0.0) for c in n.children), default=0.0)
        return dist[id(self.root)]

    def scale_edges(self, divisor):
        for n in self.nodes():
            if n.length:
                n.length /= divisor

    def to_newick(self):
        def fmt(n):
            s = ""
            if n.children:
                s = "(" + ",".join(fmt(c) for c in n.children) + ")"
            if n.label is not None:
                s += _quote(n.label)
            if n.length is not None:
                s += ":" + _num(n.length)
            return s
        return fmt(self.root) + ";"


def _num(x):
    return repr(float(x))


def _quote(label):
    return "'" + label.replace("'", "''") + "'" if re.search(r"[\s()\[\]':;,]", label) else label


def parse_newick(text):
    """Parse one newick statement (up to the first ';').  Comments in [...] are dropped."""
    s = re.sub(r"\[[^\]]*\]", "", text).strip()
    pos = 0
    n = len(s)

    def skip_ws():
        nonlocal pos
        while pos < n and s[pos].isspace():
            pos += 1

    def read_label():
        nonlocal pos
        skip_ws()
        if pos < n and s[pos] == "'":
            pos += 1
            out = []
            while pos < n:
                if s[pos] == "'":
                    if pos + 1 < n and s[pos + 1] == "'":
                        out.append("'")
                        pos += 2
                        continue
                    pos += 1
                    break
                out.append(s[pos])
                pos += 1
            return "".join(out)
        start = pos
        while pos < n and s[pos] not in ",():;" and not s[pos].isspace():
            pos += 1
        lab = s[start:pos]
        return lab.replace("_", " ") if lab else None

    def read_length():
        nonlocal pos
        skip_ws()
        if pos < n and s[pos] == ":":
            pos += 1
            skip_ws()
            start = pos
            while pos < n and s[pos] not in ",();" and not s[pos].isspace():
                pos += 1
            return float(s[start:pos])
        return None

    root = Node()
    stack = [root]
    skip_ws()
    if pos >= n:
        raise ValueError("empty tree statement")
    # iterative descent: '(' opens a child list, ',' a sibling, ')' closes
    cur = root
    if s[pos] != "(":
        cur.label = read_label()
        cur.length = read_length()
        return Tree(root)
    while pos < n:
        ch = s[pos]
        if ch == "(":
            child = Node()
            cur.children.append(child)
            stack.append(cur)
            cur = child
            pos += 1
            skip_ws()
            if pos < n and s[pos] != "(":
                cur.label = read_label()
                cur.length = read_length()
        elif ch == ",":
            parent = stack[-1]
            cur = Node()
            parent.children.append(cur)
            pos += 1
            skip_ws()
            if pos < n and s[pos] != "(":
                cur.label = read_label()
                cur.length = read_length()
        elif ch == ")":
            cur = stack.pop()
            pos += 1
            lab = read_label()
            if lab:
                cur.label = lab
            cur.length = read_length()
        elif ch == ";":
            break
        elif ch.isspace():
            pos += 1
        else:
            raise ValueError("unexpected character %r at offset %d of the tree statement" % (ch, pos))
    if len(stack) != 1:
        raise ValueError("unbalanced parentheses in the tree statement")
    return Tree(root)


def parse_nexus_tree(text):
    """First TREE statement of the first TREES block; TRANSLATE table applied to leaf labels."""
    block = re.search(r"begin\s+trees\s*;(.*?)end\s*;", text, flags=re.I | re.S)
    if not block:
        raise ValueError("no TREES block in the nexus file")
    body = block.group(1)
    translate = {}
    tr = re.search(r"translate\s+(.*?);", body, flags=re.I | re.S)
    if tr:
        for item in tr.group(1).split(","):
            parts = item.strip().split(None, 1)
            if len(parts) == 2:
                translate[parts[0]] = parts[1].strip().strip("'")
    m = re.search(r"\btree\s+[^=]+=\s*(?:\[[^\]]*\]\s*)*(.*?;)", body, flags=re.I | re.S)
    if not m:
        raise ValueError("no TREE statement in the TREES block")
    tree = parse_newick(m.group(1))
    if translate:
        for leaf in tree.leaf_nodes():
            if leaf.label in translate:
                leaf.label = translate[leaf.label]
    return tree


def read_tree(path, schema="newick"):
    with open(path) as fh:
        text = fh.read()
    schema = schema.lower()
    if schema == "newick":
        return parse_newick(text)
    if schema == "nexus":
        return parse_nexus_tree(text)
    raise ValueError("unknown tree format %r (newick or nexus)" % schema)
