/* TEST INFRASTRUCTURE ONLY -- empty autoconf header for the oracle build. */
