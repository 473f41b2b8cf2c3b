"""The Story fixtures of tests/golden/make_golden.py, rebuilt from this repo's
pieces (shared by the CPU and GPU story tests)."""

# Same-shape (4x12) chapters of the list-style story (make_golden.STORY_LIST_CHAPTERS).
LIST_CHAPTERS = (
    ('cliff_walk', None),
    ('chain_walk', ['............', '.....P......', '............', '............']),
    ('cliff_walk', ['............', '............', '........P...', '............']),
)
