"""Oracle package: CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY --
see oracle/restatement.py for the rules about who may import it."""
