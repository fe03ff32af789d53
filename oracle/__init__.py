"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU checkers for the Phenograph Jaccard edge-weighting path:

* ``binding.Oracle``      our plain-C restatement (oracle/jaccard_oracle.c)
* ``binding.Reference``   the reference's own sources compiled unmodified into
                          oracle/_ref/ (R runtime stubbed by oracle/rshim/)
* ``numpy_ref``           an independent brute-force numpy restatement (small n)
* ``louvain``             clustcells()-equivalent edge preprocessing + the
                          reference's ModularityOptimizer (oracle/_ref/modopt)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``gficf_b200`` never does.
"""
