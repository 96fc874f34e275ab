"""TEST INFRASTRUCTURE ONLY — CPU restatement of the PyGPR hot path (see pygpr_oracle.py).

Nothing under ``pygpr_b200/`` may import this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs.
"""
