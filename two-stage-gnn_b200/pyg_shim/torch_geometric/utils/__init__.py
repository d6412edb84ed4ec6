"""torch_geometric.utils: imported as a module by train*.py:7, nothing is called."""
