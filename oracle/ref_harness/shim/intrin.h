// oracle/ test infrastructure: MSVC <intrin.h> stand-in (utils.h:8 of the reference includes it)
#include <x86intrin.h>
