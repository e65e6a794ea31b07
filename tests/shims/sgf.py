"""Stand-in for the third-party `sgf` package (pypi sgf 0.5, absent from this image) — TEST INFRASTRUCTURE, used only by
tests/golden/make_golden.py so that the reference's own replay_sgf (core/eval_dataset.py:80) can run here.  It offers what
utils/sgf_wrapper.py:103-106 and eval_dataset.py touch: parse(text).children[0].root, node.properties (ident -> list of str),
node.next (main line: first variation), node.first.  The hot path never parses SGF."""


class ParseException(Exception):
    pass


class Node:
    def __init__(self, previous):
        self.properties = {}
        self.previous = previous
        self.next = None
        self.variations = []
        self.first = previous is None
        if previous is not None:
            if previous.next is None:
                previous.next = self
            else:
                previous.variations.append(self)


class GameTree:
    def __init__(self):
        self.nodes = []
        self.children = []

    @property
    def root(self):
        return self.nodes[0]


class Collection:
    def __init__(self):
        self.children = []


def _value(text, i):
    out = []
    i += 1
    while True:
        if i >= len(text):
            raise ParseException('unterminated value')
        ch = text[i]
        if ch == '\\':
            if i + 1 < len(text):
                out.append(text[i + 1])
            i += 2
        elif ch == ']':
            return ''.join(out), i + 1
        else:
            out.append(ch)
            i += 1


def _tree(text, i, previous):
    """text[i] == '(' ; returns (GameTree, index after the closing parenthesis)."""
    tree = GameTree()
    i += 1
    last = previous
    node = None
    while True:
        if i >= len(text):
            raise ParseException('unterminated tree')
        ch = text[i]
        if ch == ';':
            node = Node(last)
            last = node
            tree.nodes.append(node)
            i += 1
        elif ch == '(':
            child, i = _tree(text, i, last)
            tree.children.append(child)
        elif ch == ')':
            if not tree.nodes:
                raise ParseException('empty tree')
            return tree, i + 1
        elif ch.isalpha():
            j = i
            while text[j].isalpha():
                j += 1
            ident = ''.join(c for c in text[i:j] if c.isupper())
            i = j
            vals = []
            while True:
                while i < len(text) and text[i].isspace():
                    i += 1
                if i < len(text) and text[i] == '[':
                    v, i = _value(text, i)
                    vals.append(v)
                else:
                    break
            if node is None or not vals:
                raise ParseException('bad property')
            node.properties.setdefault(ident, []).extend(vals)
        else:
            i += 1


def parse(text):
    col = Collection()
    i = 0
    while True:
        i = text.find('(', i)
        if i < 0:
            break
        tree, i = _tree(text, i, None)
        col.children.append(tree)
    if not col.children:
        raise ParseException('no game')
    return col
