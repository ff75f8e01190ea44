// stand-in for yaml-cpp (absent from this image): deliberately empty (oracle/ref_host compiles no YAML users).
#pragma once
