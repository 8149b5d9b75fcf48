// oracle/ test infrastructure: MSVC <varargs.h> stand-in (utils.h:9 of the reference includes it)
#include <stdarg.h>
