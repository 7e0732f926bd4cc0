"""ogb -- offline stand-in for the Open Graph Benchmark loaders the reference scripts import
(main_dgl_arxiv_gat.py:10, main_dgl_molhiv_gcn.py:13-14, kernel/utils.py:5).  Datasets are seeded
synthetic graphs of the named dataset's shape (dgl.data.synthetic.SHAPES); evaluators compute the
real metrics on whatever predictions they are given."""
__version__ = "0.0+synthetic"
