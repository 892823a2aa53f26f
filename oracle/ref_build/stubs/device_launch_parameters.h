// empty stand-in (oracle/_ref host build only)
#pragma once
