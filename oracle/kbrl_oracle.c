/* placeholder until the KBRL oracle lands */ int orc_kbrl_placeholder;
