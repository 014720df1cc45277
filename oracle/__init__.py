"""TEST INFRASTRUCTURE: CPU parity oracle (see oracle/vkgs_oracle.c). Never imported by the product."""
