"""CPU oracle for the nano_gicp hot path — TEST INFRASTRUCTURE ONLY (see oracle/oracle_gicp.cpp)."""
