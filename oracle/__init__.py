"""CPU oracle for the MergeRec hot paths -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``mergerec_b200`` (the product) never does.
"""
