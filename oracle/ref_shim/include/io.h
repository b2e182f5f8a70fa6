// TEST INFRASTRUCTURE ONLY: empty stand-in for MSVC <io.h> (oracle/_ref build).
#pragma once
