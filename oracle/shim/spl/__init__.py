"""ORACLE / TEST INFRASTRUCTURE ONLY -- serial stand-in for the third-party `spl` package
(pyccel/spl, unpinned in /root/reference/requirements.txt:4, absent from the tree).  See
oracle/README.md.  Never imported by the product (poms_b200/)."""
