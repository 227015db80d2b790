"""HARNESS (not product): runs the reference's own caller code -- model/pointgroup.py, model/speaker.py -- unmodified
over d3net_b200.pointgroup_ops, with stand-ins for the third-party packages this image lacks.  See d3net_stub.py."""
