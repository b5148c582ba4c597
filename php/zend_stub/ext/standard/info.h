/* declaration-only stand-in, see ../../php.h */
#ifndef ACB200_ZEND_STUB_INFO_H
#define ACB200_ZEND_STUB_INFO_H
void php_info_print_table_start(void);
void php_info_print_table_row(int num_cols, ...);
void php_info_print_table_end(void);
#endif
