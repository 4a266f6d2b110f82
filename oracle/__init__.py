"""TEST INFRASTRUCTURE ONLY. See oracle/README.md. Importable from tests/, bench.py's cpu_baseline /
--impl reference legs and __graft_entry__.smoke() -- never from the product package."""
