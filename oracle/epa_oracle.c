/* placeholder: EPA restatement lands in a later commit */
typedef int ogjk_epa_oracle_placeholder;
