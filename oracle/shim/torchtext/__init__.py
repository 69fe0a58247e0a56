"""TEST INFRASTRUCTURE — import-only stand-in for torchtext (requirements.txt), which the reference's
datasets/scene_graph.py imports at module level (tokenizer, GloVe vocabulary).  The conversion under test
(`convert_one_gqa_scene_graph`, `query_and_translate`) touches none of it: tests build the object with
`object.__new__` and hand it a stub vocabulary.  Never imported by the product package."""
