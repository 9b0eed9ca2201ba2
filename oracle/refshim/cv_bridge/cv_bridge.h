// stand-in (TEST INFRASTRUCTURE, see oracle/refshim/README.md): the reference headers only include this file
