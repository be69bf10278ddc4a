"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference hot path (sst_oracle), the import shim that runs the unmodified
reference Python (ref_shim), the recipe that compiles the reference C++/CUDA sources (build_ref) and the golden-vector generator
(make_golden).  Nothing under sst_b200/ imports this package (tests/test_abi.py::test_product_never_imports_oracle)."""
