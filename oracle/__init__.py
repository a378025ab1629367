"""Oracle: CPU restatement of the reference hot path. TEST INFRASTRUCTURE — never imported by the product."""
