"""Host span helpers against tests/golden/spans.json (outputs of the reference's own get_actions / get_spans /
get_stats, tests/golden/make_golden_spans.py)."""
import ast
import json
import os

from cliora_b200.analysis.utils import get_actions, get_spans, get_spans_from_tree, get_stats

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'spans.json')))


def test_actions_and_spans_match_reference():
    for c in GOLD['trees']:
        actions = get_actions(c['tree'])
        assert actions == c['actions']
        assert [list(s) for s in get_spans(actions)] == c['spans']
        # the nested-tuple walker used to check the device span kernel visits the same spans in the same order
        assert [list(s) for s in get_spans_from_tree(ast.literal_eval(c['tree']))] == c['spans']


def test_stats_match_reference():
    for c in GOLD['stats']:
        a = {tuple(s) for s in c['a']}
        b = {tuple(s) for s in c['b']}
        assert list(get_stats(a, b)) == c['stats']


def test_word_tokens_with_several_characters():
    assert get_actions('((the cat) (sat down))') == [0, 0, 1, 0, 0, 1, 1]
