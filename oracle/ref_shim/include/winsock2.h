// TEST INFRASTRUCTURE ONLY: the reference's swap.h pulls ntohl/ntohs from <winsock2.h>.
#pragma once
#include <arpa/inet.h>
