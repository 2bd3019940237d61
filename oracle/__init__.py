"""Test infrastructure: CPU oracle of brille's interpolation path (see oracle/brille_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product package ``brille_b200`` never does.
"""
